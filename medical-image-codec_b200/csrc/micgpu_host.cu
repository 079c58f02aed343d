// micgpu_host.cu -- host side of libmicgpu.so: container parsing, batch planning,
// scratch management and the C ABI declared in include/micgpu.h.
//
// Mirrors the orchestration layer of the reference (parallelstrips.go:270-330,
// multiframecompress.go:227-315, multiframe.go:96-142) with the goroutine /
// pthread pools (ojph/mic_parallel.c:131-189) replaced by one batched launch
// sequence per call: K1 tables -> K2 ANS -> K3 RLE/escape -> K4 wavefront.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "host_common.h"
#include "mic_device.cuh"

using namespace micgpu;

using namespace micgpu_host;

namespace {

uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

int nstates_index(unsigned n) { return n == 1 ? 0 : n == 2 ? 1 : n == 4 ? 2 : 3; }

struct AnsPlan {
  int nstates = 0, max_log = 0, mode = 0, slots = 0, grid = 0;
  bool serial = false;   // thread-per-unit kernel (k_ans_serial.cu)
  // serial kernel: the list cut into segments whose tables fit one CTA together
  std::vector<int2> segs;             // (first list entry, entries)
  std::vector<uint32_t> slot_off;     // per list entry: byte offset inside its segment's shared memory
  size_t seg_smem = 0;
  size_t dev_off = 0;                 // byte offset of [segs | slot_off] in d_segs
};

// Streams with at most this many states go through the thread-per-unit kernel (MICGPU_K2_SERIAL_MAXN: 0 = never, 1, 2, 4, 8).
// The kernel costs ~35 cycles per SYMBOL whatever N is (issue slots of its one warp per scheduler), the lane-parallel
// kernel ~177 cycles per ROUND of N symbols: measured on the bench batch (tools/k2s_maxn.sh) 2-state 11.3 ms, 4-state
// 10.1 against 12.9 ms, 8-state 10.2 against 7.2 ms -- so N <= 4 goes here and 8-state (and rANS-8) stays lane-parallel.
int serial_max_n() {
  static const int v = [] { const char* e = getenv("MICGPU_K2_SERIAL_MAXN"); return e ? atoi(e) : 4; }();
  return v;
}

struct TemporalGroup {
  unsigned long long out_off, fpx;
  int nframes;
  int first_is_residual;   // the group is a shard that starts inside the stack (sums relative to a zero carry)
};

}  // namespace

struct micgpu_decoder {
  int device = 0;
  int sm_count = 148;
  size_t smem_optin = 227 * 1024;
  std::mutex mu;
  std::vector<MicUnit> units;
  std::vector<int> lists[4];
  std::vector<int> huff;              // canonical-Huffman units (k_huff.cu); they are in none of the ANS lists
  std::vector<int> spatial;
  std::vector<TemporalGroup> temporal;
  std::vector<std::pair<unsigned long long, unsigned long long>> zero_ranges;   // output elements no unit covers (offset, count)
  AnsPlan ans[4];
  unsigned long long sym_total = 0, tab_total = 0, d_total = 0, m_total = 0, out_need = 0;
  int max_log_all = 5, max_w = 1, max_h = 1;
  int n_grad = 0;             // spatial units coded with the gradient-adaptive predictor
  int n_rle = 0;              // RLE-kind units (temporal residual frames, wavelet coefficient streams)
  int k1_grid = 1;
  unsigned long long k1_stride = 0;
  bool committed = false;
  int launches = 0;
  DevBuf d_units, d_list, d_tabA, d_tabS, d_states, d_D, d_M, d_k1, d_comp, d_out, d_queue;
  DevBuf d_segs;            // thread-per-unit ANS kernel: segment tables
  DevBuf d_park, d_peer;    // multi-device temporal MIC2: this range's relative last frame; staged copies of the peers' frames
  DevBuf d_norm, d_psym;    // split table build: present-symbol lists (normalised count, symbol value) at each unit's tab_off
  bool k1_split = true, k1_fallback = false;
  bool all_aligned = false;   // every spatial unit of the current run: width % 8 == 0 and a 16 B aligned first pixel
  DevBuf d_jobs, d_bytes;   // MIC3: fill / blit job tables, byte-typed pixel output
  DevBuf d_wA, d_wB, d_wflags;   // WaveletV2: int32 ping-pong planes, escape flags
  MicUnit* h_units = nullptr;   // pinned staging copy
  size_t h_units_cap = 0;
  cudaStream_t stream = nullptr;  // used by the host-buffer convenience calls
  // K3 -> K4 run as PARTS independent unit ranges on their own streams: both kernels are latency bound at under half
  // of the issue slots, so the wavefront of one range fills gaps of the run expansion of the next (measured: -3 %)
  static constexpr int PARTS = 8;   // upper bound; MICGPU_PARTS picks fewer (default 4)
  cudaStream_t part_stream[PARTS] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[PARTS] = {}, ev_k3[PARTS] = {};
  // optional per-kernel timing (CUDA events on the launch stream)
  bool profiling = false;
  std::vector<cudaEvent_t> ev;
  std::vector<std::string> ev_names;
  int ev_used = 0;

  ~micgpu_decoder() {
    cudaSetDevice(device);
    d_units.release(); d_list.release(); d_tabA.release(); d_tabS.release(); d_states.release();
    d_D.release(); d_M.release(); d_k1.release(); d_norm.release(); d_psym.release(); d_park.release(); d_peer.release(); d_segs.release(); d_queue.release(); d_comp.release(); d_out.release(); d_jobs.release(); d_bytes.release(); d_wA.release(); d_wB.release(); d_wflags.release();
    if (h_units) cudaFreeHost(h_units);
    if (stream) cudaStreamDestroy(stream);
    for (int p = 0; p < PARTS; p++) {
      if (part_stream[p]) cudaStreamDestroy(part_stream[p]);
      if (ev_join[p]) cudaEventDestroy(ev_join[p]);
      if (ev_k3[p]) cudaEventDestroy(ev_k3[p]);
    }
    if (ev_fork) cudaEventDestroy(ev_fork);
    for (auto e : ev) cudaEventDestroy(e);
  }
};

namespace {

// Peek the framing of one FSE stream (fse2state.go:102-116, fsedecompressu16.go:55-61).
void peek_frame(const uint8_t* f, size_t len, MicUnit& u) {
  u.nstates = 1;
  u.rans = 0;
  u.count = 0;
  u.status = MIC_OK;
  size_t hdr = 0;
  if (len >= 2 && f[0] == 0xFF && (f[1] == 0x84 || f[1] == 0x08 || f[1] == 0x04 || f[1] == 0x02)) {
    u.nstates = f[1] == 0x02 ? 2 : f[1] == 0x04 ? 4 : 8;
    u.rans = f[1] == 0x08;
    if (len < 6) { u.status = MIC_E_HEADER; u.table_log = 5; return; }
    u.count = rd32(f + 2);
    hdr = 6;
  }
  if (len < hdr + 5 || len >= (1u << 28)) {   // ncount needs 4 bytes, the bitstream at least one
    u.status = len >= (1u << 28) ? MIC_E_UNSUPPORTED : MIC_E_HEADER;
    u.table_log = 5;
    return;
  }
  u.table_log = (f[hdr] & 0xF) + 5;
  if (u.table_log > 16) { u.status = MIC_E_NCOUNT; u.table_log = 5; }   // tablelogAbsoluteMax is 17; encoders stop at 16
}

int plan_commit(micgpu_decoder* d) {
  cudaSetDevice(d->device);
  d->sym_total = d->tab_total = d->d_total = d->m_total = 0;
  d->max_log_all = 5;
  d->max_w = d->max_h = 1;
  d->n_grad = 0;
  d->n_rle = 0;
  for (auto& l : d->lists) l.clear();
  d->huff.clear();
  d->spatial.clear();
  for (size_t i = 0; i < d->units.size(); i++) {
    MicUnit& u = d->units[i];
    const unsigned long long px = (unsigned long long)u.width * u.height;
    // symbol capacity: exact for N-state streams; for 1-state the stream is bounded by the
    // RLE worst case (every pixel escaped, plus run headers)
    const bool counted = u.nstates > 1 || u.rans == MIC_CODER_HUFF;   // the frame states its symbol count
    unsigned long long cap = counted ? u.count : 3 * px + 4096;
    if (cap > 0xFFFFFFF0ull) { u.status = MIC_E_UNSUPPORTED; cap = 0; }
    if (counted && cap > 3 * px + 4096) { u.status = MIC_E_SIZE; cap = 0; }   // count cannot exceed the RLE worst case
    u.sym_cap = (unsigned)cap;
    u.sym_off = d->sym_total;
    d->sym_total += (cap + 15) & ~15ull;
    u.tab_off = d->tab_total;
    d->tab_total += 1ull << u.table_log;
    // a Huffman unit reserves tableLog-16 regions (code table in tabA, identity in tabS) but takes no part in K1
    if (u.rans != MIC_CODER_HUFF) d->max_log_all = std::max(d->max_log_all, (int)u.table_log);
    if (u.kind == MIC_KIND_SPATIAL) {
      u.wp = (u.width + 8 + 63) & ~63u;   // rows start on 128 B lines;   // room for the 0..7 pixel row phase (mic_unit.h align0)
      u.d_off = d->d_total;
      d->d_total += (unsigned long long)u.wp * u.height;
      u.m_off = d->m_total;
      d->m_total += (unsigned long long)(u.wp >> 5) * u.height;
      d->spatial.push_back((int)i);
      if (u.predictor) d->n_grad++;
      d->max_w = std::max(d->max_w, (int)u.width);
      d->max_h = std::max(d->max_h, (int)u.height);
    } else {
      u.wp = 0; u.d_off = 0; u.m_off = 0;
      if (u.kind == MIC_KIND_RLE) d->n_rle++;
    }
    if (u.rans == MIC_CODER_HUFF) d->huff.push_back((int)i);
    else d->lists[nstates_index(u.nstates)].push_back((int)i);
    d->out_need = std::max(d->out_need, u.out_off + px);
  }
  // slots of one warp run in lockstep to the longest unit: keep similar lengths together
  for (auto& l : d->lists)
    std::stable_sort(l.begin(), l.end(), [&](int a, int b) {
      const MicUnit &ua = d->units[a], &ub = d->units[b];
      const unsigned long long ca = ua.nstates > 1 ? ua.count : ua.comp_len, cb = ub.nstates > 1 ? ub.count : ub.comp_len;
      return ca > cb;
    });
  // ---- K2 launch shapes ----------------------------------------------------
  static const int NS[4] = {1, 2, 4, 8};
  const size_t budget = d->smem_optin;
  for (int g = 0; g < 4; g++) {
    AnsPlan& a = d->ans[g];
    a.nstates = NS[g];
    const int n = (int)d->lists[g].size();
    if (!n) { a.grid = 0; continue; }
    int ml = 5;
    for (int i : d->lists[g]) ml = std::max(ml, (int)d->units[i].table_log);
    a.max_log = ml;
    const int slots_max = 32;   // one slot per warp, at most 32 warps per CTA
    const int need_per_sm = (n + d->sm_count - 1) / d->sm_count;
    auto fit = [&](int mode) {   // most slots whose tables, rings and exchange area fit one CTA
      int s = 0;
      while (s < 64 && ans_decode_smem_bytes(ml, mode, s + 1) <= budget) s++;
      return s;
    };
    a.serial = false;
    bool list_rans = false;   // rANS-8 frames share the 8-state list; only the lane-parallel kernel decodes them
    for (int i : d->lists[g]) list_rans |= d->units[i].rans == MIC_CODER_RANS;
    if (a.nstates <= serial_max_n() && !list_rans) {
      auto fit_s = [&](int mode) {
        int s = 0;
        while (s < 128 && ans_serial_smem_bytes(ml, mode, s + 1) <= budget) s++;
        return s;
      };
      const int want = std::min(need_per_sm, 128);
      // Cell format.  Split cells (u16 newState array + u8 nbBits array, 3 B per cell) put nothing between the load and
      // the chain: 10 % faster rounds than 4-byte cells (two ALU instructions to take the cell apart), 40 % faster than
      // the 2-byte form of tableLog 16 (flag word + find-leading-one).  They are taken whenever the plan's units per SM
      // fit in them, and for tableLog 16 (one unit per SM in any format); otherwise 2-byte cells, which hold 1.5x the
      // units (residency is what bounds a full batch).  MICGPU_K2S_SPLITCELLS=0: the previous choice (4-byte / 2-byte).
      static const int split_cfg = [] { const char* e = getenv("MICGPU_K2S_SPLITCELLS"); return e ? atoi(e) : 1; }();
      int smode;
      if (split_cfg == 0) smode = fit_s(0) >= want ? 0 : 1;
      else smode = fit_s(3) >= want ? 3 : ((ml == 16 && fit_s(3) >= 1) ? 3 : 1);
      const int f = fit_s(smode);
      if (f >= 1) {
        a.serial = true;
        a.mode = smode;
        // Segments: consecutive list entries (sorted by length) whose tables fit one CTA, at most `want` of them so that
        // a small batch still spreads over every SM.  Tables keep their own size: 4 KB and 8 KB tables share a CTA.
        a.segs.clear();
        a.slot_off.assign(n, 0);
        a.seg_smem = 0;
        int max_slots = 1;
        for (int i = 0; i < n;) {
          size_t used = 0;
          int cnt = 0;
          while (i + cnt < n && cnt < want) {
            const size_t ub = ans_serial_unit_bytes((int)d->units[d->lists[g][i + cnt]].table_log, smode);
            if (used + ub > budget) break;
            a.slot_off[i + cnt] = (uint32_t)used;
            used += ub;
            cnt++;
          }
          if (cnt == 0) { cnt = 1; used = ans_serial_unit_bytes((int)d->units[d->lists[g][i]].table_log, smode); }   // cannot happen: f >= 1
          a.segs.push_back(make_int2(i, cnt));
          a.seg_smem = std::max(a.seg_smem, used);
          max_slots = std::max(max_slots, cnt);
          i += cnt;
        }
        a.slots = max_slots;
        const size_t per_cta = a.seg_smem + 1024;
        const int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>((228 * 1024) / per_cta, 16));
        a.grid = std::min((int)a.segs.size(), d->sm_count * ctas_per_sm);
        continue;
      }
    }
    int mode = 0;
    // 2-byte cells hold tableLog <= 15; the packed kernel (N > 1) adds a bit array for tableLog 16
    if (fit(0) < std::min(need_per_sm, slots_max)) mode = (ml <= 15 || (ml == 16 && a.nstates > 1)) && fit(1) > fit(0) ? 1 : 0;
    if (fit(mode) < 1) mode = 2;
    a.mode = mode;
    int slots = std::min(slots_max, std::max(1, std::min(fit(mode), need_per_sm)));
    a.slots = slots;
    // CTAs per SM by shared memory (1 KB reserved per CTA) and by 2048 threads
    const size_t per_cta = ans_decode_smem_bytes(ml, mode, slots) + 1024;
    int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>((228 * 1024) / per_cta, (size_t)std::max(1, 64 / slots)));
    a.grid = std::min((n + slots - 1) / slots, d->sm_count * ctas_per_sm);
  }
  // ---- K1 scratch ------------------------------------------------------------
  const unsigned long long per_cta = 18ull << d->max_log_all;
  // one CTA per unit while they all fit in one wave (16 CTAs of 128 threads per SM): the serial ncount parse of a
  // unit is ~0.6 ms on its own, so a second wave doubles the kernel
  int g1 = std::min<int>((int)d->units.size(), d->sm_count * 16);
  const unsigned long long k1_budget = 1024ull << 20;
  if ((unsigned long long)g1 * per_cta > k1_budget) g1 = (int)std::max<unsigned long long>(d->sm_count / 2, k1_budget / per_cta);
  g1 = std::max(1, std::min<int>(g1, (int)d->units.size()));
  d->k1_grid = g1;
  d->k1_stride = (per_cta + 255) & ~255ull;

  int rc;
  const size_t nu = d->units.size();
  if ((rc = d->d_units.ensure(std::max<size_t>(nu, 1) * sizeof(MicUnit)))) return rc;
  if ((rc = d->d_list.ensure(std::max<size_t>(nu, 1) * 2 * sizeof(int)))) return rc;
  if ((rc = d->d_tabA.ensure((d->tab_total + 64) * sizeof(uint32_t)))) return rc;
  if ((rc = d->d_tabS.ensure((d->tab_total + 64) * sizeof(uint16_t)))) return rc;
  if ((rc = d->d_states.ensure((d->sym_total + 64) * sizeof(uint16_t)))) return rc;
  if ((rc = d->d_D.ensure((d->d_total + 64) * sizeof(uint16_t)))) return rc;
  if ((rc = d->d_M.ensure((d->m_total + 64) * sizeof(uint32_t)))) return rc;
  // Split table build (K1a parse + K1b build) unless MICGPU_K1_SPLIT=0; the one-kernel path stays for the units it
  // cannot take: rANS streams, tableLog > 13, more than 2048 present symbols (only possible from tableLog 12 up).
  static const bool split_cfg = [] { const char* e = getenv("MICGPU_K1_SPLIT"); return !(e && e[0] == '0'); }();
  d->k1_split = split_cfg;
  bool any_rans = false;
  for (const MicUnit& u : d->units) any_rans |= u.rans == MIC_CODER_RANS;
  d->k1_fallback = !d->k1_split || any_rans || d->max_log_all >= 12;
  if (d->k1_fallback && (rc = d->d_k1.ensure((size_t)d->k1_grid * d->k1_stride))) return rc;
  if (d->k1_split) {
    if ((rc = d->d_norm.ensure((d->tab_total + 64) * sizeof(int32_t)))) return rc;
    if ((rc = d->d_psym.ensure((d->tab_total + 64) * sizeof(uint16_t)))) return rc;
  }
  if ((rc = d->d_queue.ensure(256))) return rc;
  if (nu > d->h_units_cap) {
    if (d->h_units) cudaFreeHost(d->h_units);
    d->h_units = nullptr;
    CUDA_TRY(cudaMallocHost(&d->h_units, (nu + nu / 4 + 16) * sizeof(MicUnit)));
    d->h_units_cap = nu + nu / 4 + 16;
  }
  // unit index lists: [N=1 | N=2 | N=4 | N=8] then the spatial list, then the Huffman units (each unit is in exactly one
  // of the coder lists, so 2 * nu entries hold everything)
  std::vector<int> flat;
  flat.reserve(2 * nu);
  for (auto& l : d->lists) flat.insert(flat.end(), l.begin(), l.end());
  flat.insert(flat.end(), d->spatial.begin(), d->spatial.end());
  flat.insert(flat.end(), d->huff.begin(), d->huff.end());
  if (!flat.empty()) CUDA_TRY(cudaMemcpy(d->d_list.p, flat.data(), flat.size() * sizeof(int), cudaMemcpyHostToDevice));
  // segment tables of the thread-per-unit ANS kernel: [segs | slot offsets] per state-count group
  {
    std::vector<uint8_t> blob;
    for (int g = 0; g < 4; g++) {
      AnsPlan& a = d->ans[g];
      if (!a.serial || a.grid == 0 || d->lists[g].empty()) continue;
      a.dev_off = blob.size();
      const size_t sb = a.segs.size() * sizeof(int2), ob = a.slot_off.size() * sizeof(uint32_t);
      blob.resize(blob.size() + ((sb + ob + 15) & ~(size_t)15));
      memcpy(blob.data() + a.dev_off, a.segs.data(), sb);
      memcpy(blob.data() + a.dev_off + sb, a.slot_off.data(), ob);
    }
    if ((rc = d->d_segs.ensure(blob.size() + 64))) return rc;
    if (!blob.empty()) CUDA_TRY(cudaMemcpy(d->d_segs.p, blob.data(), blob.size(), cudaMemcpyHostToDevice));
  }
  d->committed = true;
  return 0;
}

void prof_mark(micgpu_decoder* d, const char* name, cudaStream_t st) {
  if (!d->profiling) return;
  if (d->ev_used >= (int)d->ev.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    d->ev.push_back(e);
    d->ev_names.push_back("");
  }
  d->ev_names[d->ev_used] = name;
  cudaEventRecord(d->ev[d->ev_used++], st);
}

// K4 over a slice of the spatial list: the row-scan kernel, then the wavefront kernel for the units it handed back
// (uint16 wrap: none on encoder-made streams) -- or the wavefront kernel alone with MICGPU_K4=wave / for units wider than
// one CTA can chain.  Returns the number of launches.
// gradient-predictor units of the slice (PICA strips, DecompressSingleFrameGrad): the avg kernels skip them, this one skips the rest
int launch_k4_grad(micgpu_decoder* d, MicUnit* du, const int* list, int n, void* d_out, cudaStream_t st) {
  if (!d->n_grad) return 0;
  launch_grad_wavefront(du, list, n, (const uint16_t*)d->d_D.p, (const uint32_t*)d->d_M.p, (uint16_t*)d_out, d->max_h, st);
  return 1;
}

int launch_k4(micgpu_decoder* d, MicUnit* du, const int* list, int n, void* d_out, cudaStream_t st) {
  static const bool wave_only = [] { const char* e = getenv("MICGPU_K4"); return e && e[0] == 'w'; }();
  if (n <= 0) return 0;
  if (!wave_only && launch_delta_rowscan(du, list, n, (const uint16_t*)d->d_D.p, (const uint32_t*)d->d_M.p, (uint16_t*)d_out, d->max_w, d->all_aligned, st)) {
    launch_delta_wavefront(du, list, n, (const uint16_t*)d->d_D.p, (const uint32_t*)d->d_M.p, (uint16_t*)d_out, d->max_w, d->max_h, st, 1);
    return 2 + launch_k4_grad(d, du, list, n, d_out, st);
  }
  launch_delta_wavefront(du, list, n, (const uint16_t*)d->d_D.p, (const uint32_t*)d->d_M.p, (uint16_t*)d_out, d->max_w, d->max_h, st, 0);
  return 1 + launch_k4_grad(d, du, list, n, d_out, st);
}

int run_device_locked(micgpu_decoder* d, const void* d_comp, size_t comp_bytes, void* d_out, size_t out_elems, cudaStream_t st) {
  if (!d->committed) return fail(MICGPU_E_HEADER, "decoder plan not committed");
  if (out_elems < d->out_need) return fail(MICGPU_E_SIZE, "output buffer holds %zu elements, plan needs %llu", out_elems, d->out_need);
  for (const MicUnit& u : d->units)
    if (u.comp_off + u.comp_len > comp_bytes) return fail(MICGPU_E_HEADER, "unit extends past the compressed buffer");
  CUDA_TRY(cudaSetDevice(d->device));
  d->launches = 0;
  const int nu = (int)d->units.size();
  if (!nu) return 0;
  d->all_aligned = true;
  for (MicUnit& u : d->units) {
    u.align0 = (unsigned)(((reinterpret_cast<uintptr_t>(d_out) >> 1) + u.out_off) & 7u);
    if (u.kind == MIC_KIND_SPATIAL && (u.align0 || (u.width & 7u))) d->all_aligned = false;
  }
  memcpy(d->h_units, d->units.data(), nu * sizeof(MicUnit));
  CUDA_TRY(cudaMemcpyAsync(d->d_units.p, d->h_units, nu * sizeof(MicUnit), cudaMemcpyHostToDevice, st));
  if (d->m_total) CUDA_TRY(cudaMemsetAsync(d->d_M.p, 0, d->m_total * sizeof(uint32_t), st));
  for (const auto& z : d->zero_ranges) CUDA_TRY(cudaMemsetAsync((uint16_t*)d_out + z.first, 0, z.second * sizeof(uint16_t), st));
  MicUnit* du = (MicUnit*)d->d_units.p;
  const uint8_t* comp = (const uint8_t*)d_comp;
  d->ev_used = 0;
  if (d->k1_split) {
    prof_mark(d, "k_parse_ncount", st);
    launch_parse_ncount(du, nu, comp, (int32_t*)d->d_norm.p, (uint16_t*)d->d_psym.p, st);
    prof_mark(d, "k_build_dtable", st);
    launch_build_dtable(du, nu, comp, (uint32_t*)d->d_tabA.p, (uint16_t*)d->d_tabS.p, (const int32_t*)d->d_norm.p, (const uint16_t*)d->d_psym.p,
                        (unsigned int*)d->d_queue.p + 32, (uint8_t*)d->d_k1.p, d->k1_stride, d->max_log_all,
                        d->k1_fallback ? d->k1_grid : 0, d->sm_count, st);
    d->launches += d->k1_fallback ? 3 : 2;
  } else {
    prof_mark(d, "k_build_tables", st);
    launch_build_tables(du, nu, comp, (uint32_t*)d->d_tabA.p, (uint16_t*)d->d_tabS.p, (uint8_t*)d->d_k1.p, d->k1_stride,
                        d->max_log_all, d->k1_grid, st);
    d->launches++;
  }
  const int* dl = (const int*)d->d_list.p;
  size_t loff = 0;
  for (int g = 0; g < 4; g++) {
    const int n = (int)d->lists[g].size();
    if (n) {
      const AnsPlan& a = d->ans[g];
      // names as launched: the thread-per-unit kernel, the packed kernel (N > 1 with tables in shared memory), else one unit per warp
      static const char* SERIAL[4] = {"k_ans_decode_serial<1>", "k_ans_decode_serial<2>", "k_ans_decode_serial<4>", "k_ans_decode_serial<8>"};
      static const char* PACKED[4] = {"k_ans_decode<1>", "k_ans_decode_packed<2>", "k_ans_decode_packed<4>", "k_ans_decode_packed<8>"};
      static const char* ONE[4] = {"k_ans_decode<1>", "k_ans_decode<2>", "k_ans_decode<4>", "k_ans_decode<8>"};
      prof_mark(d, a.serial ? SERIAL[g] : (a.mode != 2 ? PACKED[g] : ONE[g]), st);
      if (a.serial) {
        const int2* segs = (const int2*)((const uint8_t*)d->d_segs.p + a.dev_off);
        const uint32_t* soff = (const uint32_t*)((const uint8_t*)segs + a.segs.size() * sizeof(int2));
        launch_ans_decode_serial(du, dl + loff, segs, (int)a.segs.size(), soff, a.nstates, comp, (const uint32_t*)d->d_tabA.p,
                                 (uint16_t*)d->d_states.p, a.mode, a.slots, a.seg_smem, a.grid, st);
      } else
        launch_ans_decode(du, dl + loff, n, a.nstates, comp, (const uint32_t*)d->d_tabA.p, (uint16_t*)d->d_states.p, a.max_log,
                          a.mode, a.slots, a.grid, st);
      d->launches++;
    }
    loff += n;
  }
  if (!d->huff.empty()) {
    static const int huff_serial = [] { const char* e = getenv("MICGPU_HUFF_SERIAL"); return e ? atoi(e) : 0; }();
    prof_mark(d, "k_huff_decode", st);
    launch_huff_decode(du, dl + loff + d->spatial.size(), (int)d->huff.size(), comp, (uint32_t*)d->d_tabA.p, (uint16_t*)d->d_tabS.p,
                       (uint16_t*)d->d_states.p, huff_serial, d->sm_count, st);
    d->launches++;
  }
  // K3 launch shape (k_rle.cu): small tables mean 8-bit planes (MIC3 tiles), whose run headers come every <= 124 symbols:
  // the header walk of one warp is the bound there, and CTAs of 128 threads / 2048-element chunks put 2.4x the walkers
  // on an SM (measured on 4096 tiles: K3 1.86 -> 1.45 ms; strips lose 10 % with it, profiles/README.md)
  // Plans with RLE-kind units (residual frames, wavelet streams) take the kernel that can walk a window's run headers in
  // parallel (shape 5; chosen per window from the header density: residual frames have a header every ~5 symbols and
  // their K3 time falls from 224 to 44 ms per stack, wavelet streams and spatial frames stay on the one-warp walk).
  // Tile planes measured slightly slower with it (1.53 against 1.44 ms per 4096 tiles) and keep shape 2.
  const int k3_shape = d->max_log_all <= 12 ? 2 : (d->n_rle ? 5 : 0);
  // CTAs of K3 per SM when unit ranges overlap on their own streams: K3 is a persistent queue kernel, so fewer CTAs leave
  // registers for the K4 CTAs of the previous range to co-reside (MICGPU_K3_CPS for A/B runs)
  static const bool k3_chain = [] { const char* e = getenv("MICGPU_K3_CHAIN"); return e && e[0] == '1'; }();
  static const int k3_cps = [] { const char* e = getenv("MICGPU_K3_CPS"); int v = e ? atoi(e) : 8; return v < 1 ? 1 : (v > 16 ? 16 : v); }();
  static const int parts_cfg = [] { const char* e = getenv("MICGPU_PARTS"); int v = e ? atoi(e) : 4; return v < 1 ? 1 : (v > micgpu_decoder::PARTS ? micgpu_decoder::PARTS : v); }();
  const int parts = (d->profiling || nu < 256) ? 1 : parts_cfg;   // per-kernel timing needs one stream
  if (parts == 1) {
    prof_mark(d, "k_rle_expand", st);
    launch_rle_expand(du, 0, nu, (const uint16_t*)d->d_states.p, (const uint16_t*)d->d_tabS.p, (uint16_t*)d->d_D.p,
                      (uint32_t*)d->d_M.p, (uint16_t*)d_out, d->max_log_all, std::min(nu, d->sm_count * 8), (unsigned int*)d->d_queue.p, st, k3_shape);
    d->launches++;
    if (!d->spatial.empty()) {
      prof_mark(d, "k_delta_rowscan", st);
      d->launches += launch_k4(d, du, dl + loff, (int)d->spatial.size(), d_out, st);
    }
  } else {
    if (!d->ev_fork) {
      CUDA_TRY(cudaEventCreateWithFlags(&d->ev_fork, cudaEventDisableTiming));
      for (int p = 0; p < micgpu_decoder::PARTS; p++) {
        CUDA_TRY(cudaStreamCreateWithFlags(&d->part_stream[p], cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&d->ev_join[p], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&d->ev_k3[p], cudaEventDisableTiming));
      }
    }
    CUDA_TRY(cudaEventRecord(d->ev_fork, st));
    for (int p = 0; p < parts; p++) {
      const int u0 = (int)((long long)nu * p / parts), u1 = (int)((long long)nu * (p + 1) / parts);
      cudaStream_t ps = d->part_stream[p];
      CUDA_TRY(cudaStreamWaitEvent(ps, d->ev_fork, 0));
      // MICGPU_K3_CHAIN=1: the K3 kernels of the ranges run one after the other, so K3 of range p+1 shares the SMs with K4
      // of range p instead of with the other K3 kernels
      if (k3_chain && p > 0) CUDA_TRY(cudaStreamWaitEvent(ps, d->ev_k3[p - 1], 0));
      launch_rle_expand(du, u0, u1 - u0, (const uint16_t*)d->d_states.p, (const uint16_t*)d->d_tabS.p, (uint16_t*)d->d_D.p,
                        (uint32_t*)d->d_M.p, (uint16_t*)d_out, d->max_log_all, std::min(u1 - u0, d->sm_count * k3_cps),
                        (unsigned int*)d->d_queue.p + p, ps, k3_shape);
      d->launches++;
      if (k3_chain) CUDA_TRY(cudaEventRecord(d->ev_k3[p], ps));
      // the spatial list is sorted by unit index: this range's units are one contiguous slice of it
      const int s0 = (int)(std::lower_bound(d->spatial.begin(), d->spatial.end(), u0) - d->spatial.begin());
      const int s1 = (int)(std::lower_bound(d->spatial.begin(), d->spatial.end(), u1) - d->spatial.begin());
      if (s1 > s0) {
        d->launches += launch_k4(d, du, dl + loff + s0, s1 - s0, d_out, ps);
      }
      CUDA_TRY(cudaEventRecord(d->ev_join[p], ps));
      CUDA_TRY(cudaStreamWaitEvent(st, d->ev_join[p], 0));
    }
  }
  for (const TemporalGroup& t : d->temporal) {
    prof_mark(d, "k_temporal_accumulate", st);
    launch_temporal_accumulate((uint16_t*)d_out + t.out_off, t.fpx, t.nframes, t.first_is_residual, d->sm_count, st);
    d->launches++;
  }
  prof_mark(d, "end", st);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int unit_status_locked(micgpu_decoder* d, int* status, int n, cudaStream_t st) {
  const int nu = (int)d->units.size();
  if (!nu) return 0;
  CUDA_TRY(cudaSetDevice(d->device));
  CUDA_TRY(cudaMemcpyAsync(d->h_units, d->d_units.p, nu * sizeof(MicUnit), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  int first = 0;
  for (int i = 0; i < nu; i++) {
    if (status && i < n) status[i] = d->h_units[i].status;
    if (!first && d->h_units[i].status) first = d->h_units[i].status;
  }
  if (first) fail(first, "unit decode failed (status %d)", first);
  return first;
}

int add_unit_locked(micgpu_decoder* d, const uint8_t* frame, size_t len, uint64_t comp_off, int kind, uint32_t w, uint32_t h,
                    uint64_t out_off) {
  MicUnit u;
  memset(&u, 0, sizeof u);
  u.comp_off = comp_off;
  u.comp_len = (unsigned)len;
  u.kind = (unsigned)kind;
  u.width = w;
  u.height = h;
  u.out_off = out_off;
  peek_frame(frame, len, u);
  if (w == 0 || h == 0) u.status = MIC_E_HEADER;
  d->units.push_back(u);
  d->committed = false;
  return (int)d->units.size() - 1;
}

// The fixed head of a canonical-Huffman stream (canhuffmancompressu16.go:119-128, MSB first): u32 symbol count, u16
// maxValue, u8 maxCodeLength, u16 list size.  There is no magic: the caller says that the stream is Huffman-coded.
int add_huff_unit_locked(micgpu_decoder* d, const uint8_t* s, size_t len, uint64_t comp_off, int kind, uint32_t w, uint32_t h,
                         uint64_t out_off) {
  MicUnit u;
  memset(&u, 0, sizeof u);
  u.comp_off = comp_off;
  u.comp_len = (unsigned)len;
  u.kind = (unsigned)kind;
  u.width = w;
  u.height = h;
  u.out_off = out_off;
  u.nstates = 1;
  u.rans = MIC_CODER_HUFF;
  u.table_log = 16;
  u.status = MIC_OK;
  if (len < 9 || w == 0 || h == 0) u.status = MIC_E_HEADER;
  else if (len >= (1u << 28)) u.status = MIC_E_UNSUPPORTED;
  else {
    u.count = ((uint32_t)s[0] << 24) | ((uint32_t)s[1] << 16) | ((uint32_t)s[2] << 8) | s[3];
    const unsigned max_value = ((unsigned)s[4] << 8) | s[5], max_len = s[6], nl = ((unsigned)s[7] << 8) | s[8];
    unsigned depth = 0, len_bits = 0;
    while (depth < 16 && (max_value >> depth)) depth++;
    while (len_bits < 8 && (max_len >> len_bits)) len_bits++;
    // 2^maxCodeLength table cells live in the unit's tabA region (2^16): the reference's own encoder stops at 14 bits
    // before it adds the delimiter (canhuffmancompressu16.go:168-186)
    if (max_len > 16) u.status = MIC_E_UNSUPPORTED;
    else if (72ull + (unsigned long long)nl * (depth + len_bits) > (unsigned long long)len * 8) u.status = MIC_E_HEADER;
  }
  d->units.push_back(u);
  d->committed = false;
  return (int)d->units.size() - 1;
}

struct PicsHeader {
  int w, h, nstrips, strip_h;
  size_t header_size;
};

// parallelstrips.go:270-290
int parse_pics(const uint8_t* p, size_t len, PicsHeader& ph) {
  if (len < 20 || memcmp(p, "PICS", 4) != 0) return fail(MICGPU_E_HEADER, "parallelstrips: invalid magic");
  ph.w = (int)rd32(p + 4); ph.h = (int)rd32(p + 8); ph.nstrips = (int)rd32(p + 12); ph.strip_h = (int)rd32(p + 16);
  if (ph.w <= 0 || ph.h <= 0 || ph.nstrips <= 0 || ph.strip_h <= 0) return fail(MICGPU_E_HEADER, "parallelstrips: invalid dimensions");
  if ((size_t)ph.nstrips > (len - 20) / 8) return fail(MICGPU_E_HEADER, "parallelstrips: truncated header");
  ph.header_size = 20 + (size_t)ph.nstrips * 8;
  if ((long long)(ph.nstrips - 1) * ph.strip_h >= ph.h) return fail(MICGPU_E_HEADER, "parallelstrips: strip table exceeds image height");
  return 0;
}

int add_pics_locked(micgpu_decoder* d, const uint8_t* pics, size_t len, uint64_t comp_off, uint64_t out_off, int* w, int* h) {
  PicsHeader ph;
  int rc = parse_pics(pics, len, ph);
  if (rc) return rc;
  for (int s = 0; s < ph.nstrips; s++) {
    const size_t so = rd32(pics + 20 + (size_t)s * 8), sl = rd32(pics + 24 + (size_t)s * 8);
    const size_t start = ph.header_size + so, end = start + sl;
    if (end > len || start > end) return fail(MICGPU_E_HEADER, "strip %d: offset out of bounds", s);
    const int y0 = s * ph.strip_h;
    const int sh = std::min(ph.strip_h, ph.h - y0);
    add_unit_locked(d, pics + start, sl, comp_off + start, MIC_KIND_SPATIAL, (uint32_t)ph.w, (uint32_t)sh,
                    out_off + (uint64_t)y0 * ph.w);
  }
  // A strip table that stops short of the image leaves the remaining rows as make() left them: zero
  // (parallelstrips.go:288).  The output buffer here is recycled scratch, so those rows are cleared explicitly.
  const unsigned long long covered = std::min<unsigned long long>((unsigned long long)ph.nstrips * ph.strip_h, (unsigned long long)ph.h);
  if (covered < (unsigned long long)ph.h)
    d->zero_ranges.push_back({out_off + covered * ph.w, ((unsigned long long)ph.h - covered) * ph.w});
  d->out_need = std::max<unsigned long long>(d->out_need, out_off + (unsigned long long)ph.w * ph.h);
  if (w) *w = ph.w;
  if (h) *h = ph.h;
  return 0;
}

// PICA (parallelstripsadaptive.go:143-212): "PICA" w h nStrips | nStrips x {y0, offset, length, flags} u32 | blobs.
// Strip s covers rows [y0_s, y0_{s+1}) (the last one runs to h); flags bit 0 = gradient-adaptive predictor.
struct PicaHeader {
  int w, h, nstrips;
  size_t header_size;
};

int parse_pica(const uint8_t* p, size_t len, PicaHeader& ph) {
  if (len < 16 || memcmp(p, "PICA", 4) != 0) return fail(MICGPU_E_HEADER, "pica: invalid magic");
  ph.w = (int)rd32(p + 4); ph.h = (int)rd32(p + 8); ph.nstrips = (int)rd32(p + 12);
  if (ph.nstrips < 0 || (size_t)ph.nstrips > (len - 16) / 16) return fail(MICGPU_E_HEADER, "pica: truncated header");
  if (ph.w <= 0 || ph.h <= 0 || ph.nstrips <= 0) return fail(MICGPU_E_HEADER, "pica: invalid dimensions");
  ph.header_size = 16 + (size_t)ph.nstrips * 16;
  return 0;
}

int add_pica_locked(micgpu_decoder* d, const uint8_t* pica, size_t len, uint64_t comp_off, uint64_t out_off, int* w, int* h) {
  PicaHeader ph;
  int rc = parse_pica(pica, len, ph);
  if (rc) return rc;
  // validate the whole strip table before any unit is added (a half-populated plan must not survive an error)
  for (int s = 0; s < ph.nstrips; s++) {
    const uint8_t* e = pica + 16 + (size_t)s * 16;
    const uint64_t y0 = rd32(e), so = rd32(e + 4), sl = rd32(e + 8);
    const uint64_t y1 = s + 1 < ph.nstrips ? (uint64_t)rd32(e + 16) : (uint64_t)ph.h;
    const uint64_t start = ph.header_size + so, end = start + sl;
    if (end > len) return fail(MICGPU_E_HEADER, "strip %d: offset out of bounds", s);
    // the reference slices out[y0*w:] and sizes the strip as y1 - y0 rows: anything else panics there
    if (y0 >= y1 || y1 > (uint64_t)ph.h) return fail(MICGPU_E_HEADER, "pica: strip %d covers rows [%llu, %llu) of %d", s, (unsigned long long)y0, (unsigned long long)y1, ph.h);
  }
  for (int s = 0; s < ph.nstrips; s++) {
    const uint8_t* e = pica + 16 + (size_t)s * 16;
    const uint32_t y0 = rd32(e), so = rd32(e + 4), sl = rd32(e + 8), fl = rd32(e + 12);
    const uint32_t y1 = s + 1 < ph.nstrips ? rd32(e + 16) : (uint32_t)ph.h;
    const size_t start = ph.header_size + so;
    const int ui = add_unit_locked(d, pica + start, sl, comp_off + start, MIC_KIND_SPATIAL, (uint32_t)ph.w, y1 - y0, out_off + (uint64_t)y0 * ph.w);
    d->units[ui].predictor = fl & 1u;
  }
  // rows above the first strip stay as make() left them (zero)
  const uint32_t first_y0 = rd32(pica + 16);
  if (first_y0 > 0) d->zero_ranges.push_back({out_off, (unsigned long long)first_y0 * ph.w});
  d->out_need = std::max<unsigned long long>(d->out_need, out_off + (unsigned long long)ph.w * ph.h);
  if (w) *w = ph.w;
  if (h) *h = ph.h;
  return 0;
}

struct Mic2Header {
  int w, h, n, temporal;
  size_t data_off;
};

// multiframe.go:96-127
int parse_mic2(const uint8_t* p, size_t len, Mic2Header& mh) {
  if (len < 20) return fail(MICGPU_E_HEADER, "MIC2: file too small");
  if (memcmp(p, "MIC2", 4) != 0) return fail(MICGPU_E_HEADER, "MIC2: invalid magic");
  mh.w = (int)rd32(p + 4); mh.h = (int)rd32(p + 8); mh.n = (int)rd32(p + 12);
  mh.temporal = (p[16] & 0x02) != 0;
  if (mh.w <= 0 || mh.h <= 0 || mh.n < 0) return fail(MICGPU_E_HEADER, "MIC2: invalid dimensions");
  if ((size_t)mh.n > (len - 20) / 8) return fail(MICGPU_E_HEADER, "MIC2: file truncated in frame table");
  mh.data_off = 20 + (size_t)mh.n * 8;
  return 0;
}

int add_mic2_locked(micgpu_decoder* d, const uint8_t* p, size_t len, uint64_t comp_off, uint64_t out_off, int last_frame,
                    Mic2Header& mh, int first_frame = 0) {
  int rc = parse_mic2(p, len, mh);
  if (rc) return rc;
  const unsigned long long fpx = (unsigned long long)mh.w * mh.h;
  const int nf = last_frame < 0 ? mh.n : (int)std::min<long long>(mh.n, (long long)last_frame + 1);
  if (first_frame < 0 || first_frame > nf) return fail(MICGPU_E_HEADER, "MIC2: frame range starts at %d of %d", first_frame, nf);
  if ((unsigned long long)mh.w * mh.h > 0xFFFFFFFFull) return fail(MICGPU_E_UNSUPPORTED, "MIC2: frames of %dx%d pixels", mh.w, mh.h);
  out_off -= (uint64_t)first_frame * fpx;   // frame i lands at out_off + (i - first_frame) * fpx
  for (int i = first_frame; i < nf; i++) {
    const size_t o = rd32(p + 20 + (size_t)i * 8), l = rd32(p + 24 + (size_t)i * 8);
    if (o > len - mh.data_off || l > len - mh.data_off - o) return fail(MICGPU_E_HEADER, "MIC2: frame %d data extends beyond file", i);
    const bool residual = mh.temporal && i > 0;
    // residual frames expand to exactly w*h ZigZag words (multiframecompress.go:165-175)
    const int ui = add_unit_locked(d, p + mh.data_off + o, l, comp_off + mh.data_off + o, residual ? MIC_KIND_RLE : MIC_KIND_SPATIAL,
                                   residual ? (uint32_t)fpx : (uint32_t)mh.w, residual ? 1u : (uint32_t)mh.h, out_off + (uint64_t)i * fpx);
    if (residual) d->units[ui].exact_len = 1;   // a shorter residual frame would leave stale scratch in the running sum
  }
  const int count = nf - first_frame;
  if (mh.temporal && (count > 1 || (count == 1 && first_frame > 0)))
    d->temporal.push_back(TemporalGroup{out_off + (uint64_t)first_frame * fpx, fpx, count, first_frame > 0 ? 1 : 0});
  return 0;
}

// ---- default per-device contexts for the one-shot calls ---------------------
std::mutex g_mu;
micgpu_decoder* g_default[64] = {nullptr};
constexpr int PIPE_DEPTH = 6;
std::mutex g_pipe_mu;
micgpu_decoder* g_pipe[64][PIPE_DEPTH] = {};

// Devices the one-call batch entry points spread their units over (micgpu_init); empty = the current device only.
std::mutex g_dev_mu;
std::vector<int> g_devices;
std::vector<int> configured_devices() {
  std::lock_guard<std::mutex> lk(g_dev_mu);
  return g_devices;
}

// C++ twin of shard.partition_by_bytes: `parts` contiguous ranges of n units with near-equal byte sums;
// every unit owned exactly once, trailing ranges may be empty when there are fewer units than parts.
void partition_by_bytes(const uint64_t* sizes, uint64_t n, int parts, std::vector<uint64_t>& cuts) {
  cuts.assign((size_t)parts + 1, n);
  unsigned __int128 total = 0;
  for (uint64_t i = 0; i < n; i++) total += sizes[i];
  uint64_t lo = 0;
  unsigned __int128 acc = 0;
  for (int r = 0; r < parts; r++) {
    cuts[r] = lo;
    uint64_t hi = lo;
    // target = total * (r + 1) / parts, compared without the division
    while (hi < n && ((acc + sizes[hi]) * (unsigned)parts <= total * (unsigned)(r + 1) || (hi == lo && n - hi > (uint64_t)(parts - 1 - r)))) {
      acc += sizes[hi];
      hi++;
    }
    if (r == parts - 1) hi = n;
    lo = hi;
  }
  cuts[parts] = n;
}

// Contexts of the multi-device calls, one per entry of the device list (entries may name the same device).
std::mutex g_multi_mu;
std::vector<micgpu_decoder*> g_multi;
micgpu_decoder* multi_decoder(int dev, int slot) {
  std::lock_guard<std::mutex> lk(g_multi_mu);
  if (slot < 0 || slot >= 64) return nullptr;
  if ((int)g_multi.size() <= slot) g_multi.resize(slot + 1, nullptr);
  if (g_multi[slot] && g_multi[slot]->device != dev) { delete g_multi[slot]; g_multi[slot] = nullptr; }
  if (!g_multi[slot]) g_multi[slot] = micgpu_decoder_create(dev);
  return g_multi[slot];
}

micgpu_decoder* default_decoder(int dev) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (dev < 0 || dev >= 64) return nullptr;
  if (!g_default[dev]) g_default[dev] = micgpu_decoder_create(dev);
  return g_default[dev];
}

// host comp -> device, run, device -> host
int run_host_locked(micgpu_decoder* d, const uint8_t* comp, size_t comp_bytes, uint16_t* out, size_t out_elems) {
  CUDA_TRY(cudaSetDevice(d->device));
  int rc;
  if ((rc = d->d_comp.ensure(comp_bytes + 256))) return rc;
  if ((rc = d->d_out.ensure(std::max<size_t>(out_elems, 1) * sizeof(uint16_t)))) return rc;
  CUDA_TRY(cudaMemcpyAsync(d->d_comp.p, comp, comp_bytes, cudaMemcpyHostToDevice, d->stream));
  if ((rc = run_device_locked(d, d->d_comp.p, comp_bytes, d->d_out.p, out_elems, d->stream))) return rc;
  CUDA_TRY(cudaMemcpyAsync(out, d->d_out.p, out_elems * sizeof(uint16_t), cudaMemcpyDeviceToHost, d->stream));
  return unit_status_locked(d, nullptr, 0, d->stream);
}

}  // namespace

// ============================ C ABI ============================================
extern "C" {

int micgpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

const char* micgpu_last_error(void) { return err_slot().c_str(); }

void* micgpu_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) {
    fail(MICGPU_E_ALLOC, "cudaMallocHost(%zu) failed", bytes);
    return nullptr;
  }
  return p;
}
void micgpu_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int micgpu_init(const int* devices, int n) {
  int have = 0;
  if (cudaGetDeviceCount(&have) != cudaSuccess || have <= 0) return fail(MICGPU_E_CUDA, "no CUDA device available (libmicgpu has no CPU fallback)");
  std::vector<int> v;
  if (!devices || n <= 0) {
    for (int i = 0; i < have; i++) v.push_back(i);     // all visible devices
  } else {
    for (int i = 0; i < n; i++) {
      if (devices[i] < 0 || devices[i] >= have) return fail(MICGPU_E_CUDA, "device %d out of range [0,%d)", devices[i], have);
      v.push_back(devices[i]);   // a device may appear more than once: that many host threads and contexts share it
    }
  }
  if (v.size() > 32) return fail(MICGPU_E_UNSUPPORTED, "at most 32 device-list entries");
  std::lock_guard<std::mutex> lk(g_dev_mu);
  g_devices = v;
  return (int)v.size();
}

void micgpu_shutdown(void) {
  {
    std::lock_guard<std::mutex> lk(g_dev_mu);
    g_devices.clear();
  }
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& d : g_default) {
    delete d;
    d = nullptr;
  }
  {
    std::lock_guard<std::mutex> lk3(g_multi_mu);
    for (auto& d : g_multi) delete d;
    g_multi.clear();
  }
  std::lock_guard<std::mutex> lk2(g_pipe_mu);
  for (auto& row : g_pipe)
    for (auto& d : row) {
      delete d;
      d = nullptr;
    }
}

micgpu_decoder* micgpu_decoder_create(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    fail(MICGPU_E_CUDA, "no CUDA device available (libmicgpu has no CPU fallback)");
    return nullptr;
  }
  if (device < 0 || device >= n) {
    fail(MICGPU_E_CUDA, "device %d out of range [0,%d)", device, n);
    return nullptr;
  }
  if (cudaSetDevice(device) != cudaSuccess) {
    fail(MICGPU_E_CUDA, "cudaSetDevice(%d) failed", device);
    return nullptr;
  }
  micgpu_decoder* d = new micgpu_decoder();
  d->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) {
    d->sm_count = prop.multiProcessorCount;
    d->smem_optin = prop.sharedMemPerBlockOptin;
  }
  if (cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking) != cudaSuccess) {
    fail(MICGPU_E_CUDA, "cudaStreamCreate failed");
    delete d;
    return nullptr;
  }
  return d;
}

void micgpu_decoder_destroy(micgpu_decoder* d) { delete d; }

int micgpu_decoder_begin(micgpu_decoder* d) {
  if (!d) return fail(MICGPU_E_HEADER, "null decoder");
  std::lock_guard<std::mutex> lk(d->mu);
  d->units.clear();
  d->temporal.clear();
  d->zero_ranges.clear();
  d->out_need = 0;
  d->committed = false;
  return 0;
}

int micgpu_decoder_add_unit(micgpu_decoder* d, const uint8_t* frame, size_t frame_len, uint64_t comp_off, int kind,
                            uint32_t width, uint32_t height, uint64_t out_off) {
  if (!d || !frame) return fail(MICGPU_E_HEADER, "null argument");
  if (kind != MICGPU_KIND_SPATIAL && kind != MICGPU_KIND_RLE) return fail(MICGPU_E_HEADER, "unknown unit kind %d", kind);
  std::lock_guard<std::mutex> lk(d->mu);
  return add_unit_locked(d, frame, frame_len, comp_off, kind, width, height, out_off);
}

int micgpu_decoder_add_huff_unit(micgpu_decoder* d, const uint8_t* stream, size_t len, uint64_t comp_off, int kind,
                                 uint32_t width, uint32_t height, uint64_t out_off) {
  if (!d || !stream) return fail(MICGPU_E_HEADER, "null argument");
  if (kind != MICGPU_KIND_SPATIAL && kind != MICGPU_KIND_RLE) return fail(MICGPU_E_HEADER, "unknown unit kind %d", kind);
  std::lock_guard<std::mutex> lk(d->mu);
  return add_huff_unit_locked(d, stream, len, comp_off, kind, width, height, out_off);
}

int micgpu_decoder_add_pics(micgpu_decoder* d, const uint8_t* pics, size_t len, uint64_t comp_off, uint64_t out_off,
                            int* width, int* height) {
  if (!d || !pics) return fail(MICGPU_E_HEADER, "null argument");
  std::lock_guard<std::mutex> lk(d->mu);
  return add_pics_locked(d, pics, len, comp_off, out_off, width, height);
}

int micgpu_decoder_add_pica(micgpu_decoder* d, const uint8_t* pica, size_t len, uint64_t comp_off, uint64_t out_off,
                            int* width, int* height) {
  if (!d || !pica) return fail(MICGPU_E_HEADER, "null argument");
  std::lock_guard<std::mutex> lk(d->mu);
  return add_pica_locked(d, pica, len, comp_off, out_off, width, height);
}

int micgpu_decoder_add_mic2(micgpu_decoder* d, const uint8_t* mic2, size_t len, uint64_t comp_off, uint64_t out_off,
                            int* width, int* height, int* frames, int* temporal) {
  if (!d || !mic2) return fail(MICGPU_E_HEADER, "null argument");
  std::lock_guard<std::mutex> lk(d->mu);
  Mic2Header mh;
  int rc = add_mic2_locked(d, mic2, len, comp_off, out_off, -1, mh);
  if (rc) return rc;
  if (width) *width = mh.w;
  if (height) *height = mh.h;
  if (frames) *frames = mh.n;
  if (temporal) *temporal = mh.temporal;
  return 0;
}

int micgpu_decoder_add_mic2_range(micgpu_decoder* d, const uint8_t* mic2, size_t len, uint64_t comp_off, uint64_t out_off,
                                  int first_frame, int frame_count, int* width, int* height, int* frames, int* temporal) {
  if (!d || !mic2) return fail(MICGPU_E_HEADER, "null argument");
  if (first_frame < 0 || frame_count < 0) return fail(MICGPU_E_HEADER, "MIC2: negative frame range");
  std::lock_guard<std::mutex> lk(d->mu);
  Mic2Header mh;
  int rc = parse_mic2(mic2, len, mh);
  if (rc) return rc;
  // validated in 64 bits before any unit is added, so a failed call leaves the plan as it was
  if ((long long)first_frame + frame_count > mh.n) return fail(MICGPU_E_HEADER, "MIC2: frame range %d+%d exceeds %d frames", first_frame, frame_count, mh.n);
  const size_t units_before = d->units.size(), temporal_before = d->temporal.size();
  rc = add_mic2_locked(d, mic2, len, comp_off, out_off, first_frame + frame_count - 1, mh, first_frame);
  if (rc) { d->units.resize(units_before); d->temporal.resize(temporal_before); return rc; }
  if (width) *width = mh.w;
  if (height) *height = mh.h;
  if (frames) *frames = mh.n;
  if (temporal) *temporal = mh.temporal;
  return 0;
}

int micgpu_temporal_add_carry(void* d_frames, const void* d_carry, uint64_t frame_px, int nframes, void* cuda_stream) {
  if (!d_frames || !d_carry) return fail(MICGPU_E_HEADER, "null argument");
  int dev = 0, sms = 148;
  CUDA_TRY(cudaGetDevice(&dev));
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  launch_temporal_add_carry((uint16_t*)d_frames, (const uint16_t*)d_carry, frame_px, nframes, sms, (cudaStream_t)cuda_stream);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int micgpu_temporal_add_carry_peers(void* d_frames, const void* const* d_peer_last, int npeers, uint64_t frame_px, int nframes,
                                    void* cuda_stream) {
  if (!d_frames || (npeers > 0 && !d_peer_last)) return fail(MICGPU_E_HEADER, "null argument");
  if (npeers < 0 || npeers > 16) return fail(MICGPU_E_HEADER, "at most 16 earlier ranges (got %d)", npeers);
  if (npeers == 0) return 0;
  int dev = 0, sms = 148;
  CUDA_TRY(cudaGetDevice(&dev));
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  PeerFrames pf;
  pf.n = npeers;
  pf.aligned = (reinterpret_cast<uintptr_t>(d_frames) % 16 == 0) && (frame_px % 8 == 0);
  for (int q = 0; q < npeers; q++) {
    if (!d_peer_last[q]) return fail(MICGPU_E_HEADER, "null peer frame %d", q);
    pf.last[q] = (const uint16_t*)d_peer_last[q];
    if (reinterpret_cast<uintptr_t>(d_peer_last[q]) % 16) pf.aligned = 0;
  }
  launch_temporal_add_carry_peers((uint16_t*)d_frames, pf, frame_px, nframes, sms, (cudaStream_t)cuda_stream);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

void* micgpu_device_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) {
    fail(MICGPU_E_ALLOC, "cudaMalloc(%zu) failed", bytes);
    return nullptr;
  }
  return p;
}

void micgpu_device_free(void* p) {
  if (p) cudaFree(p);
}

int micgpu_ipc_export(const void* d_ptr, void* handle64) {
  if (!d_ptr || !handle64) return fail(MICGPU_E_HEADER, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, const_cast<void*>(d_ptr)));
  memcpy(handle64, &h, 64);
  return 0;
}

int micgpu_ipc_open(const void* handle64, void** d_ptr) {
  if (!handle64 || !d_ptr) return fail(MICGPU_E_HEADER, "null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  CUDA_TRY(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int micgpu_ipc_close(void* d_ptr) {
  if (!d_ptr) return 0;
  CUDA_TRY(cudaIpcCloseMemHandle(d_ptr));
  return 0;
}

int micgpu_decoder_commit(micgpu_decoder* d) {
  if (!d) return fail(MICGPU_E_HEADER, "null decoder");
  std::lock_guard<std::mutex> lk(d->mu);
  return plan_commit(d);
}

int micgpu_decoder_unit_count(const micgpu_decoder* d) { return d ? (int)d->units.size() : 0; }

int micgpu_decoder_run_device(micgpu_decoder* d, const void* d_comp, size_t comp_bytes, void* d_out, size_t out_elems,
                              void* cuda_stream) {
  if (!d) return fail(MICGPU_E_HEADER, "null decoder");
  std::lock_guard<std::mutex> lk(d->mu);
  return run_device_locked(d, d_comp, comp_bytes, d_out, out_elems, (cudaStream_t)cuda_stream);
}

int micgpu_decoder_unit_status(micgpu_decoder* d, int* status, int n, void* cuda_stream) {
  if (!d) return fail(MICGPU_E_HEADER, "null decoder");
  std::lock_guard<std::mutex> lk(d->mu);
  return unit_status_locked(d, status, n, (cudaStream_t)cuda_stream);
}

int micgpu_decoder_last_launches(const micgpu_decoder* d) { return d ? d->launches : 0; }

int micgpu_decoder_set_profiling(micgpu_decoder* d, int on) {
  if (!d) return fail(MICGPU_E_HEADER, "null decoder");
  std::lock_guard<std::mutex> lk(d->mu);
  d->profiling = on != 0;
  return 0;
}

int micgpu_decoder_kernel_times(micgpu_decoder* d, char* names, size_t names_cap, float* ms, int cap) {
  if (!d) return fail(MICGPU_E_HEADER, "null decoder");
  std::lock_guard<std::mutex> lk(d->mu);
  const int n = d->ev_used > 0 ? d->ev_used - 1 : 0;
  if (n == 0) return 0;
  CUDA_TRY(cudaEventSynchronize(d->ev[d->ev_used - 1]));
  std::string joined;
  for (int i = 0; i < n && i < cap; i++) {
    float t = 0;
    CUDA_TRY(cudaEventElapsedTime(&t, d->ev[i], d->ev[i + 1]));
    ms[i] = t;
    if (i) joined += ";";
    joined += d->ev_names[i];
  }
  if (names && names_cap) {
    strncpy(names, joined.c_str(), names_cap - 1);
    names[names_cap - 1] = 0;
  }
  return std::min(n, cap);
}

int micgpu_decoder_run_host(micgpu_decoder* d, const uint8_t* comp, size_t comp_bytes, uint16_t* out, size_t out_elems) {
  if (!d) return fail(MICGPU_E_HEADER, "null decoder");
  std::lock_guard<std::mutex> lk(d->mu);
  return run_host_locked(d, comp, comp_bytes, out, out_elems);
}

// ---- one-shot container calls --------------------------------------------------
}  // extern "C"

namespace {

// One in-flight PICS chunk on one decoder: everything is enqueued on d->stream, nothing is waited for.
struct PicsPending {
  int i0 = 0, i1 = 0;
  std::vector<int> first_unit, hdr_rc;
  bool active = false;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // MICGPU_TRACE: start, H2D done, kernels done, D2H done
  double t_enq0 = 0, t_enq1 = 0;
  ~PicsPending() {
    for (auto& e : ev)
      if (e) cudaEventDestroy(e);
  }
};

bool trace_on() {
  static const bool v = [] { const char* e = getenv("MICGPU_TRACE"); return e && e[0] == '1'; }();
  return v;
}
double host_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}
void trace_mark(PicsPending& P, int k, cudaStream_t st) {
  if (!trace_on()) return;
  if (!P.ev[k]) cudaEventCreate(&P.ev[k]);
  cudaEventRecord(P.ev[k], st);
}

int pics_enqueue(micgpu_decoder* d, int i0, int i1, const uint8_t* const* blobs, const size_t* lens, uint16_t* const* outs,
                 const size_t* caps, PicsPending& P) {
  const int n = i1 - i0;
  P.t_enq0 = trace_on() ? host_ms() : 0;
  d->units.clear();
  d->temporal.clear();
  d->zero_ranges.clear();
  d->out_need = 0;
  // layout: blobs back to back (64-byte aligned) in one device buffer; outputs back to back
  std::vector<uint64_t> coff(n), ooff(n + 1);
  P.i0 = i0; P.i1 = i1;
  P.first_unit.assign(n + 1, 0);
  P.hdr_rc.assign(n, 0);
  uint64_t ctot = 0, otot = 0;
  for (int k = 0; k < n; k++) {
    const int i = i0 + k;
    coff[k] = ctot;
    ooff[k] = otot;
    P.first_unit[k] = (int)d->units.size();
    int w = 0, h = 0;
    const size_t zr_before = d->zero_ranges.size();
    const unsigned long long need_before = d->out_need;
    P.hdr_rc[k] = add_pics_locked(d, blobs[i], lens[i], ctot, otot, &w, &h);
    if (!P.hdr_rc[k] && (size_t)w * h > caps[i]) P.hdr_rc[k] = fail(MICGPU_E_SIZE, "image %d: output buffer too small", i);
    if (!P.hdr_rc[k]) otot += (uint64_t)w * h;
    else { d->units.resize(P.first_unit[k]); d->zero_ranges.resize(zr_before); d->out_need = need_before; }
    ctot += (lens[i] + 63) & ~(size_t)63;
  }
  ooff[n] = otot;
  P.first_unit[n] = (int)d->units.size();
  int rc = plan_commit(d);
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(d->device));
  if ((rc = d->d_comp.ensure(ctot + 256))) return rc;
  if ((rc = d->d_out.ensure(std::max<uint64_t>(otot, 1) * sizeof(uint16_t)))) return rc;
  // copies are merged when neighbouring images are also neighbours in host memory (a caller that keeps a batch in one
  // buffer gets one H2D and one D2H per chunk instead of one per image)
  trace_mark(P, 0, d->stream);
  for (int k = 0; k < n;) {
    if (P.hdr_rc[k]) { k++; continue; }
    int j = k;
    size_t bytes = lens[i0 + k];
    while (j + 1 < n && !P.hdr_rc[j + 1] && blobs[i0 + j + 1] == blobs[i0 + j] + (coff[j + 1] - coff[j])) { j++; bytes = (size_t)(coff[j] - coff[k]) + lens[i0 + j]; }
    CUDA_TRY(cudaMemcpyAsync((uint8_t*)d->d_comp.p + coff[k], blobs[i0 + k], bytes, cudaMemcpyHostToDevice, d->stream));
    k = j + 1;
  }
  trace_mark(P, 1, d->stream);
  if ((rc = run_device_locked(d, d->d_comp.p, ctot, d->d_out.p, otot, d->stream))) return rc;
  trace_mark(P, 2, d->stream);
  for (int k = 0; k < n;) {
    if (P.hdr_rc[k]) { k++; continue; }
    int j = k;
    while (j + 1 < n && !P.hdr_rc[j + 1] && outs[i0 + j + 1] == outs[i0 + j] + (ooff[j + 1] - ooff[j])) j++;
    CUDA_TRY(cudaMemcpyAsync(outs[i0 + k], (uint16_t*)d->d_out.p + ooff[k], (ooff[j + 1] - ooff[k]) * sizeof(uint16_t), cudaMemcpyDeviceToHost, d->stream));
    k = j + 1;
  }
  if (!d->units.empty())
    CUDA_TRY(cudaMemcpyAsync(d->h_units, d->d_units.p, d->units.size() * sizeof(MicUnit), cudaMemcpyDeviceToHost, d->stream));
  trace_mark(P, 3, d->stream);
  P.t_enq1 = trace_on() ? host_ms() : 0;
  P.active = true;
  return 0;
}

// Wait for the chunk and fold unit statuses into per-image statuses (first failing strip wins, parallelstrips.go:324-328)
int pics_finish(micgpu_decoder* d, PicsPending& P, int* status, int* first) {
  if (!P.active) return 0;
  P.active = false;
  CUDA_TRY(cudaSetDevice(d->device));
  const double t_w0 = trace_on() ? host_ms() : 0;
  CUDA_TRY(cudaStreamSynchronize(d->stream));
  if (trace_on() && P.ev[0]) {
    float h2d = 0, ker = 0, d2h = 0;
    cudaEventElapsedTime(&h2d, P.ev[0], P.ev[1]);
    cudaEventElapsedTime(&ker, P.ev[1], P.ev[2]);
    cudaEventElapsedTime(&d2h, P.ev[2], P.ev[3]);
    fprintf(stderr, "[micgpu trace] images %d-%d: host enqueue %.2f ms (at %.2f), device h2d %.2f kernels %.2f d2h %.2f ms, host waited %.2f ms (until %.2f)\n",
            P.i0, P.i1, P.t_enq1 - P.t_enq0, P.t_enq0, h2d, ker, d2h, host_ms() - t_w0, host_ms());
  }
  const int n = P.i1 - P.i0;
  for (int k = 0; k < n; k++) {
    int st = P.hdr_rc[k];
    for (int u = P.first_unit[k]; !st && u < P.first_unit[k + 1]; u++) st = d->h_units[u].status;
    if (status) status[P.i0 + k] = st;
    if (!*first && st) { *first = st; if (st != P.hdr_rc[k]) fail(st, "image %d: strip decode failed (status %d)", P.i0 + k, st); }
  }
  return 0;
}

// Extra decoder contexts so that the H2D copy of chunk c+1, the kernels of chunk c and the D2H copy of chunk c-1 overlap.

micgpu_decoder* pipe_decoder(int dev, int k) {
  std::lock_guard<std::mutex> lk(g_pipe_mu);
  if (dev < 0 || dev >= 64) return nullptr;
  if (!g_pipe[dev][k]) g_pipe[dev][k] = micgpu_decoder_create(dev);
  return g_pipe[dev][k];
}

}  // namespace

extern "C" {

}  // extern "C"

namespace {

// One device's share of a PICS batch (all of it without micgpu_init).
int pics_batch_on_device(int dev, int n, const uint8_t* const* blobs, const size_t* lens, uint16_t* const* outs,
                         const size_t* caps, int* status) {
  if (n <= 0) return 0;
  if (cudaSetDevice(dev) != cudaSuccess) return fail(MICGPU_E_CUDA, "cudaSetDevice(%d) failed", dev);
  int first = 0, rc;
  if (n < 16) {
    micgpu_decoder* d = default_decoder(dev);
    if (!d) return MICGPU_E_CUDA;
    std::lock_guard<std::mutex> lk(d->mu);
    PicsPending P;
    if ((rc = pics_enqueue(d, 0, n, blobs, lens, outs, caps, P))) { cudaStreamSynchronize(d->stream); return rc; }
    if ((rc = pics_finish(d, P, status, &first))) { cudaStreamSynchronize(d->stream); return rc; }
    return first;
  }
  // Pipelined: ~16 chunks over PIPE_DEPTH decoder contexts (each with its own stream and scratch).  The ANS stage of a
  // chunk takes the same ~7.6 ms whether it holds 100 or 2000 strips (it is bound by the serial chain of one strip), and
  // a chunk's D2H copy is shorter than that, so several chunks must be computing at once to keep the PCIe link busy.
  // The first chunks are small and grow geometrically: the first D2H copy cannot start before one ANS pass (~8 ms)
  // has finished, whatever the chunk holds, so a small head chunk gets the PCIe link busy ~6 ms earlier.
  const int chunk = std::max(4, (n + 15) / 16);
  micgpu_decoder* D[PIPE_DEPTH];
  PicsPending P[PIPE_DEPTH];
  for (int k = 0; k < PIPE_DEPTH; k++)
    if (!(D[k] = pipe_decoder(dev, k))) return MICGPU_E_CUDA;
  std::unique_lock<std::mutex> locks[PIPE_DEPTH];
  for (int k = 0; k < PIPE_DEPTH; k++) locks[k] = std::unique_lock<std::mutex>(D[k]->mu);
  // An error in one chunk must not leave kernels and device->host copies of the other contexts running into the
  // caller's buffers after the call has returned (their locks and the PicsPending array die with this frame).
  auto drain = [&]() {
    for (int k = 0; k < PIPE_DEPTH; k++) {
      cudaSetDevice(D[k]->device);
      cudaStreamSynchronize(D[k]->stream);
      P[k].active = false;
    }
  };
  int c = 0, size = std::max(2, chunk / 4);
  for (int i0 = 0; i0 < n; c++) {
    const int k = c % PIPE_DEPTH;
    const int i1 = std::min(n, i0 + size);
    if ((rc = pics_finish(D[k], P[k], status, &first))) { drain(); return rc; }
    if ((rc = pics_enqueue(D[k], i0, i1, blobs, lens, outs, caps, P[k]))) { drain(); return rc; }
    i0 = i1;
    size = std::min(chunk, size * 2);
  }
  for (int k = 0; k < PIPE_DEPTH; k++)
    if ((rc = pics_finish(D[k], P[k], status, &first))) { drain(); return rc; }
  return first;
}

}  // namespace

extern "C" {

int micgpu_pics_decompress_batch(int n, const uint8_t* const* blobs, const size_t* lens, uint16_t* const* outs,
                                 const size_t* caps, int* status) {
  if (n <= 0) return 0;
  if (!blobs || !lens || !outs || !caps) return fail(MICGPU_E_HEADER, "null argument");
  const std::vector<int> devs = configured_devices();
  if (devs.size() <= 1 || n < 2 * (int)devs.size())
    return pics_batch_on_device(devs.empty() ? current_device() : devs[0], n, blobs, lens, outs, caps, status);
  // several devices: contiguous image ranges balanced by compressed bytes, one host thread per device (the goroutine
  // pool of parallelstrips.go:292-321 becomes one batched launch sequence per device); nothing is exchanged
  std::vector<uint64_t> sizes(n), cuts;
  for (int i = 0; i < n; i++) sizes[i] = lens[i];
  partition_by_bytes(sizes.data(), (uint64_t)n, (int)devs.size(), cuts);
  std::vector<int> rcs(devs.size(), 0);
  std::vector<std::string> msgs(devs.size());
  std::vector<std::thread> th;
  for (size_t k = 0; k < devs.size(); k++)
    th.emplace_back([&, k] {
      const int i0 = (int)cuts[k], i1 = (int)cuts[k + 1];
      rcs[k] = pics_batch_on_device(devs[k], i1 - i0, blobs + i0, lens + i0, outs + i0, caps + i0, status ? status + i0 : nullptr);
      if (rcs[k]) msgs[k] = err_slot();
    });
  for (auto& t : th) t.join();
  for (size_t k = 0; k < devs.size(); k++)
    if (rcs[k]) { err_slot() = msgs[k]; return rcs[k]; }   // first failing image in batch order (parallelstrips.go:324-328)
  return 0;
}

int micgpu_pics_decompress(const uint8_t* pics, size_t len, uint16_t* pixels_out, size_t cap_px, int* width, int* height) {
  if (!pics || !pixels_out) return fail(MICGPU_E_HEADER, "null argument");
  PicsHeader ph;
  int rc = parse_pics(pics, len, ph);
  if (rc) return rc;
  if (width) *width = ph.w;
  if (height) *height = ph.h;
  return micgpu_pics_decompress_batch(1, &pics, &len, &pixels_out, &cap_px, nullptr);
}

int micgpu_decompress_single_frame(const uint8_t* frame, size_t len, uint16_t* pixels_out, int width, int height) {
  if (!frame || !pixels_out || width <= 0 || height <= 0) return fail(MICGPU_E_HEADER, "bad argument");
  micgpu_decoder* d = default_decoder(current_device());
  if (!d) return MICGPU_E_CUDA;
  std::lock_guard<std::mutex> lk(d->mu);
  d->units.clear();
  d->temporal.clear();
  d->zero_ranges.clear();
  d->out_need = 0;
  add_unit_locked(d, frame, len, 0, MIC_KIND_SPATIAL, (uint32_t)width, (uint32_t)height, 0);
  int rc = plan_commit(d);
  if (rc) return rc;
  return run_host_locked(d, frame, len, pixels_out, (size_t)width * height);
}

// DecompressParallelStripsAdaptive (parallelstripsadaptive.go:143-212)
int micgpu_pica_decompress(const uint8_t* pica, size_t len, uint16_t* pixels_out, size_t cap_px, int* width, int* height) {
  if (!pica || !pixels_out) return fail(MICGPU_E_HEADER, "null argument");
  PicaHeader ph;
  int rc = parse_pica(pica, len, ph);
  if (rc) return rc;
  if (width) *width = ph.w;
  if (height) *height = ph.h;
  if ((unsigned long long)ph.w * ph.h > cap_px) return fail(MICGPU_E_SIZE, "output buffer holds %zu pixels, image has %llu", cap_px, (unsigned long long)ph.w * ph.h);
  micgpu_decoder* d = default_decoder(current_device());
  if (!d) return MICGPU_E_CUDA;
  std::lock_guard<std::mutex> lk(d->mu);
  d->units.clear();
  d->temporal.clear();
  d->zero_ranges.clear();
  d->out_need = 0;
  if ((rc = add_pica_locked(d, pica, len, 0, 0, nullptr, nullptr))) return rc;
  if ((rc = plan_commit(d))) return rc;
  return run_host_locked(d, pica, len, pixels_out, (size_t)ph.w * ph.h);
}

// DecompressSingleFrameGrad (multiframecompress.go:129-142)
int micgpu_decompress_single_frame_grad(const uint8_t* frame, size_t len, uint16_t* pixels_out, int width, int height) {
  if (!frame || !pixels_out || width <= 0 || height <= 0) return fail(MICGPU_E_HEADER, "bad argument");
  micgpu_decoder* d = default_decoder(current_device());
  if (!d) return MICGPU_E_CUDA;
  std::lock_guard<std::mutex> lk(d->mu);
  d->units.clear();
  d->temporal.clear();
  d->zero_ranges.clear();
  d->out_need = 0;
  const int ui = add_unit_locked(d, frame, len, 0, MIC_KIND_SPATIAL, (uint32_t)width, (uint32_t)height, 0);
  d->units[ui].predictor = 1;
  int rc = plan_commit(d);
  if (rc) return rc;
  return run_host_locked(d, frame, len, pixels_out, (size_t)width * height);
}

// CanHuffmanDecompressU16.Init + ReadTable + Decompress (canhuffmandecompressu16.go:31-108): the symbols of one stream
int micgpu_huff_decompress(const uint8_t* stream, size_t len, uint16_t* symbols_out, size_t cap, size_t* n_out) {
  if (!stream || !symbols_out || !n_out) return fail(MICGPU_E_HEADER, "null argument");
  if (len < 9) return fail(MICGPU_E_HEADER, "huffman: stream shorter than its fixed header");
  const size_t n = ((size_t)stream[0] << 24) | ((size_t)stream[1] << 16) | ((size_t)stream[2] << 8) | stream[3];
  *n_out = n;
  if (n > cap) return fail(MICGPU_E_SIZE, "output buffer holds %zu symbols, stream has %zu", cap, n);
  if (n == 0) return 0;
  micgpu_decoder* d = default_decoder(current_device());
  if (!d) return MICGPU_E_CUDA;
  std::lock_guard<std::mutex> lk(d->mu);
  d->units.clear();
  d->temporal.clear();
  d->zero_ranges.clear();
  d->out_need = 0;
  add_huff_unit_locked(d, stream, len, 0, MIC_KIND_RAW, (uint32_t)n, 1, 0);
  int rc = plan_commit(d);
  if (rc) return rc;
  return run_host_locked(d, stream, len, symbols_out, n);
}

// DeltaRleHuffDecompressU16.Decompress (deltarlehuffdecompressu16.go:19-39): Huffman -> RLE -> avg(top,left) predictor
int micgpu_delta_rle_huff_decompress(const uint8_t* stream, size_t len, uint16_t* pixels_out, int width, int height) {
  if (!stream || !pixels_out || width <= 0 || height <= 0) return fail(MICGPU_E_HEADER, "bad argument");
  micgpu_decoder* d = default_decoder(current_device());
  if (!d) return MICGPU_E_CUDA;
  std::lock_guard<std::mutex> lk(d->mu);
  d->units.clear();
  d->temporal.clear();
  d->zero_ranges.clear();
  d->out_need = 0;
  add_huff_unit_locked(d, stream, len, 0, MIC_KIND_SPATIAL, (uint32_t)width, (uint32_t)height, 0);
  int rc = plan_commit(d);
  if (rc) return rc;
  return run_host_locked(d, stream, len, pixels_out, (size_t)width * height);
}

static int mic2_decode(const uint8_t* mic2, size_t len, int last_frame, bool only_last, uint16_t* out, size_t cap_px, int* width,
                       int* height, int* frames, int* temporal) {
  if (!mic2 || !out) return fail(MICGPU_E_HEADER, "null argument");
  micgpu_decoder* d = default_decoder(current_device());
  if (!d) return MICGPU_E_CUDA;
  std::lock_guard<std::mutex> lk(d->mu);
  d->units.clear();
  d->temporal.clear();
  d->zero_ranges.clear();
  d->out_need = 0;
  Mic2Header mh;
  int rc = parse_mic2(mic2, len, mh);
  if (rc) return rc;
  if (width) *width = mh.w;
  if (height) *height = mh.h;
  if (frames) *frames = mh.n;
  if (temporal) *temporal = mh.temporal;
  if (only_last && (last_frame < 0 || last_frame >= mh.n)) return fail(MICGPU_E_HEADER, "frame index %d out of range [0, %d)", last_frame, mh.n);
  const size_t fpx = (size_t)mh.w * mh.h;
  size_t nout;
  if (only_last && !mh.temporal) {
    // independent mode: decode just the requested frame (multiframecompress.go:277-288)
    const size_t o = rd32(mic2 + 20 + (size_t)last_frame * 8), l = rd32(mic2 + 24 + (size_t)last_frame * 8);
    if (o > len - mh.data_off || l > len - mh.data_off - o) return fail(MICGPU_E_HEADER, "MIC2: frame %d data extends beyond file", last_frame);
    add_unit_locked(d, mic2 + mh.data_off + o, l, mh.data_off + o, MIC_KIND_SPATIAL, (uint32_t)mh.w, (uint32_t)mh.h, 0);
    nout = 1;
  } else {
    if ((rc = add_mic2_locked(d, mic2, len, 0, 0, only_last ? last_frame : -1, mh))) return rc;
    nout = only_last ? (size_t)last_frame + 1 : (size_t)mh.n;
  }
  if ((only_last ? fpx : nout * fpx) > cap_px) return fail(MICGPU_E_SIZE, "output buffer too small");
  if ((rc = plan_commit(d))) return rc;
  if (nout == 0) return 0;
  CUDA_TRY(cudaSetDevice(d->device));
  if ((rc = d->d_comp.ensure(len + 256))) return rc;
  if ((rc = d->d_out.ensure(nout * fpx * sizeof(uint16_t)))) return rc;
  CUDA_TRY(cudaMemcpyAsync(d->d_comp.p, mic2, len, cudaMemcpyHostToDevice, d->stream));
  if ((rc = run_device_locked(d, d->d_comp.p, len, d->d_out.p, nout * fpx, d->stream))) return rc;
  if (only_last)
    CUDA_TRY(cudaMemcpyAsync(out, (uint16_t*)d->d_out.p + (nout - 1) * fpx, fpx * sizeof(uint16_t), cudaMemcpyDeviceToHost, d->stream));
  else
    CUDA_TRY(cudaMemcpyAsync(out, d->d_out.p, nout * fpx * sizeof(uint16_t), cudaMemcpyDeviceToHost, d->stream));
  return unit_status_locked(d, nullptr, 0, d->stream);
}

// DecompressMultiFrame over the configured devices: frame ranges balanced by compressed bytes (multiframe.go:72-78), one
// host thread per device.  Independent mode needs nothing else.  Temporal mode (multiframecompress.go:236-258): every
// range decodes to running sums relative to a zero carry, parks its last frame, and after a host barrier ONE kernel per
// device reads the parked frames of the earlier devices straight out of their memory (peer access over NVLink; a
// staged peer copy where the topology has none), forms the carry and finishes its frames -- the single exchange step of
// the whole path.
static int mic2_decode_multi(const std::vector<int>& devs, const uint8_t* mic2, size_t len, const Mic2Header& mh0, uint16_t* out) {
  const int nd = (int)devs.size();
  const size_t fpx = (size_t)mh0.w * mh0.h;
  std::vector<uint64_t> sizes(mh0.n), cuts;
  for (int i = 0; i < mh0.n; i++) sizes[i] = rd32(mic2 + 24 + (size_t)i * 8);
  partition_by_bytes(sizes.data(), (uint64_t)mh0.n, nd, cuts);
  std::vector<micgpu_decoder*> D(nd);
  for (int k = 0; k < nd; k++)
    if (!(D[k] = multi_decoder(devs[k], k))) return MICGPU_E_CUDA;
  std::vector<std::unique_lock<std::mutex>> locks;
  for (int k = 0; k < nd; k++) locks.emplace_back(D[k]->mu);   // device-list order: no two calls can deadlock
  std::vector<int> rcs(nd, 0);
  std::vector<std::string> msgs(nd);
  auto phase = [&](auto fn) {
    std::vector<std::thread> th;
    for (int k = 0; k < nd; k++)
      th.emplace_back([&, k] {
        rcs[k] = fn(k);
        if (rcs[k]) msgs[k] = err_slot();
      });
    for (auto& t : th) t.join();
    for (int k = 0; k < nd; k++)
      if (rcs[k]) { err_slot() = msgs[k]; return rcs[k]; }
    return 0;
  };
  // ---- phase 1: every device decodes its frame range ---------------------------------------------------------
  int rc = phase([&](int k) -> int {
    const int f0 = (int)cuts[k], f1 = (int)cuts[k + 1];
    if (f1 <= f0) return 0;
    micgpu_decoder* d = D[k];
    CUDA_TRY(cudaSetDevice(d->device));
    d->units.clear();
    d->temporal.clear();
    d->zero_ranges.clear();
    d->out_need = 0;
    // only the bytes of this range travel: [span0, span1) of the container sits at offset 0 of the device buffer
    const size_t o0 = rd32(mic2 + 20 + (size_t)f0 * 8), o1 = rd32(mic2 + 20 + (size_t)(f1 - 1) * 8), l1 = rd32(mic2 + 24 + (size_t)(f1 - 1) * 8);
    size_t span0 = mh0.data_off + o0, span1 = mh0.data_off + o1 + l1;
    for (int i = f0; i < f1; i++) {   // frame tables are written in order, but nothing forces them to be
      const size_t o = rd32(mic2 + 20 + (size_t)i * 8), l = rd32(mic2 + 24 + (size_t)i * 8);
      span0 = std::min(span0, mh0.data_off + o);
      span1 = std::max(span1, mh0.data_off + o + l);
    }
    if (span1 > len) return fail(MICGPU_E_HEADER, "MIC2: frame data extends beyond file");
    Mic2Header mh;
    int r = add_mic2_locked(d, mic2, len, (uint64_t)0 - (uint64_t)span0, 0, f1 - 1, mh, f0);
    if (r) return r;
    if ((r = plan_commit(d))) return r;
    const size_t nfr = (size_t)(f1 - f0);
    if ((r = d->d_comp.ensure(span1 - span0 + 256))) return r;
    if ((r = d->d_out.ensure(nfr * fpx * sizeof(uint16_t)))) return r;
    CUDA_TRY(cudaMemcpyAsync(d->d_comp.p, mic2 + span0, span1 - span0, cudaMemcpyHostToDevice, d->stream));
    if ((r = run_device_locked(d, d->d_comp.p, span1 - span0, d->d_out.p, nfr * fpx, d->stream))) return r;
    if (mh0.temporal) {
      if ((r = d->d_park.ensure(fpx * sizeof(uint16_t) + 16))) return r;
      CUDA_TRY(cudaMemcpyAsync(d->d_park.p, (uint16_t*)d->d_out.p + (nfr - 1) * fpx, fpx * sizeof(uint16_t), cudaMemcpyDeviceToDevice, d->stream));
    } else {
      CUDA_TRY(cudaMemcpyAsync(out + (size_t)f0 * fpx, d->d_out.p, nfr * fpx * sizeof(uint16_t), cudaMemcpyDeviceToHost, d->stream));
    }
    return unit_status_locked(d, nullptr, 0, d->stream);   // synchronises the stream
  });
  if (rc || !mh0.temporal) return rc;
  // ---- phase 2 (temporal): carry from the parked frames of the earlier devices, then the copy back ----------------
  return phase([&](int k) -> int {
    const int f0 = (int)cuts[k], f1 = (int)cuts[k + 1];
    if (f1 <= f0) return 0;
    micgpu_decoder* d = D[k];
    CUDA_TRY(cudaSetDevice(d->device));
    const size_t nfr = (size_t)(f1 - f0);
    PeerFrames pf;
    pf.n = 0;
    pf.aligned = (fpx % 8 == 0) ? 1 : 0;
    int staged = 0;
    for (int q = 0; q < k; q++) {
      if (cuts[q + 1] <= cuts[q]) continue;     // an empty range carries nothing
      if (pf.n >= 16) return fail(MICGPU_E_UNSUPPORTED, "temporal MIC2 over more than 17 devices");
      const uint16_t* src = (const uint16_t*)D[q]->d_park.p;
      int can = 0;
      if (D[q]->device != d->device) {
        cudaDeviceCanAccessPeer(&can, d->device, D[q]->device);
        if (can) {
          const cudaError_t e = cudaDeviceEnablePeerAccess(D[q]->device, 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can = 0;
          cudaGetLastError();
        }
        if (!can) {   // no peer mapping on this topology: the frame is copied over instead of read in place
          int r = d->d_peer.ensure((size_t)16 * (fpx * sizeof(uint16_t) + 16));
          if (r) return r;
          uint16_t* dst = (uint16_t*)((uint8_t*)d->d_peer.p + (size_t)staged * (fpx * sizeof(uint16_t) + 16));
          CUDA_TRY(cudaMemcpyPeerAsync(dst, d->device, src, D[q]->device, fpx * sizeof(uint16_t), d->stream));
          src = dst;
          staged++;
        }
      }
      if (reinterpret_cast<uintptr_t>(src) % 16) pf.aligned = 0;
      pf.last[pf.n++] = src;
    }
    if (reinterpret_cast<uintptr_t>(d->d_out.p) % 16) pf.aligned = 0;
    if (pf.n) launch_temporal_add_carry_peers((uint16_t*)d->d_out.p, pf, fpx, (int)nfr, d->sm_count, d->stream);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out + (size_t)f0 * fpx, d->d_out.p, nfr * fpx * sizeof(uint16_t), cudaMemcpyDeviceToHost, d->stream));
    CUDA_TRY(cudaStreamSynchronize(d->stream));
    return 0;
  });
}

int micgpu_mic2_decompress(const uint8_t* mic2, size_t len, uint16_t* frames_out, size_t cap_px, int* width, int* height,
                           int* frames, int* temporal) {
  const std::vector<int> devs = configured_devices();
  if (devs.size() > 1 && mic2 && frames_out) {
    Mic2Header mh;
    int rc = parse_mic2(mic2, len, mh);
    if (rc) return rc;
    if (mh.n >= 2 * (int)devs.size() && (unsigned long long)mh.w * mh.h <= 0xFFFFFFFFull) {
      if (width) *width = mh.w;
      if (height) *height = mh.h;
      if (frames) *frames = mh.n;
      if (temporal) *temporal = mh.temporal;
      if ((size_t)mh.n * mh.w * mh.h > cap_px) return fail(MICGPU_E_SIZE, "output buffer too small");
      return mic2_decode_multi(devs, mic2, len, mh, frames_out);
    }
  }
  return mic2_decode(mic2, len, -1, false, frames_out, cap_px, width, height, frames, temporal);
}

int micgpu_mic2_decompress_frame(const uint8_t* mic2, size_t len, int frame_idx, uint16_t* pixels_out, size_t cap_px, int* width,
                                 int* height) {
  return mic2_decode(mic2, len, frame_idx, true, pixels_out, cap_px, width, height, nullptr, nullptr);
}

}  // extern "C"

// ---- MIC3 / RGB ---------------------------------------------------------------------
namespace {

struct Mic3Header {
  int width, height, tile_w, tile_h, channels, bps, ct, nlv;
  uint64_t total_tiles;
  size_t lv_off, table_off, data_off;
};
struct Mic3Level { int w, h, tx, ty, first; };

uint64_t rd64(const uint8_t* p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }

// ReadMIC3Header (wsiformat.go:169-227)
int parse_mic3(const uint8_t* p, size_t len, Mic3Header& h) {
  if (len < 48) return fail(MICGPU_E_HEADER, "MIC3: file too small");
  if (memcmp(p, "MIC3", 4) != 0) return fail(MICGPU_E_HEADER, "MIC3: invalid magic");
  if (rd32(p + 4) != 1) return fail(MICGPU_E_HEADER, "MIC3: unsupported version %u", rd32(p + 4));
  h.width = (int)rd32(p + 8); h.height = (int)rd32(p + 12); h.tile_w = (int)rd32(p + 16); h.tile_h = (int)rd32(p + 20);
  h.channels = p[24] | (p[25] << 8); h.bps = p[26]; h.ct = (p[27] & 0x02) != 0;
  h.nlv = p[28] | (p[29] << 8);
  h.total_tiles = rd64(p + 32);
  h.lv_off = 48;
  if (len < h.lv_off + (size_t)h.nlv * 20) return fail(MICGPU_E_HEADER, "MIC3: truncated level descriptors");
  h.table_off = h.lv_off + (size_t)h.nlv * 20;
  if (h.total_tiles > (len - h.table_off) / 16) return fail(MICGPU_E_HEADER, "MIC3: truncated tile offset table");
  h.data_off = h.table_off + (size_t)h.total_tiles * 16;
  if (h.tile_w <= 0 || h.tile_h <= 0 || h.width <= 0 || h.height <= 0) return fail(MICGPU_E_HEADER, "MIC3: invalid dimensions");
  if (h.tile_w > 32768 || h.tile_h > 32768) return fail(MICGPU_E_UNSUPPORTED, "MIC3: tile size %dx%d", h.tile_w, h.tile_h);
  if (h.nlv > MICGPU_WSI_MAX_LEVELS) return fail(MICGPU_E_UNSUPPORTED, "MIC3: %d pyramid levels (at most %d)", h.nlv, MICGPU_WSI_MAX_LEVELS);
  // Level descriptors are file data: every field a later index computation relies on is checked here (the Go reader
  // would panic on an out-of-range slice; a GPU kernel would write out of bounds).
  for (int l = 0; l < h.nlv; l++) {
    const uint8_t* q = p + h.lv_off + (size_t)l * 20;
    const uint32_t w = rd32(q), hh = rd32(q + 4), tx = rd32(q + 8), ty = rd32(q + 12), first = rd32(q + 16);
    if (w == 0 || hh == 0 || w > 0x7FFFFFFFu || hh > 0x7FFFFFFFu) return fail(MICGPU_E_HEADER, "MIC3: level %d has invalid dimensions", l);
    const uint64_t etx = ((uint64_t)w + h.tile_w - 1) / h.tile_w, ety = ((uint64_t)hh + h.tile_h - 1) / h.tile_h;
    if (tx != etx || ty != ety) return fail(MICGPU_E_HEADER, "MIC3: level %d tile grid %ux%u does not match %ux%u pixels", l, tx, ty, w, hh);
    if ((uint64_t)first + etx * ety > h.total_tiles) return fail(MICGPU_E_HEADER, "MIC3: level %d tiles exceed the tile table", l);
  }
  return 0;
}
Mic3Level mic3_level(const uint8_t* p, const Mic3Header& h, int l) {
  const uint8_t* q = p + h.lv_off + (size_t)l * 20;
  return Mic3Level{(int)rd32(q), (int)rd32(q + 4), (int)rd32(q + 8), (int)rd32(q + 12), (int)rd32(q + 16)};
}
// ExtractTileBlob (wsiformat.go:229-241)
int mic3_tile_blob(const uint8_t* p, size_t len, const Mic3Header& h, long long idx, const uint8_t** blob, size_t* bl) {
  if (idx < 0 || (uint64_t)idx >= h.total_tiles) return fail(MICGPU_E_HEADER, "MIC3: tile index %lld out of range", idx);
  const uint64_t o = rd64(p + h.table_off + (size_t)idx * 16), n = rd64(p + h.table_off + (size_t)idx * 16 + 8);
  const uint64_t room = (uint64_t)len - h.data_off;   // data_off <= len (parse_mic3); no sum of file fields can wrap
  if (o > room || n > room - o) return fail(MICGPU_E_HEADER, "MIC3: tile %lld data extends beyond file", idx);
  *blob = p + h.data_off + o;
  *bl = (size_t)n;
  return 0;
}

struct Blit { unsigned sx, sy, w, h; unsigned long long dst_off; unsigned dst_pitch; };
struct TileReq {
  const uint8_t* blob; size_t len;       // host copy of the tile blob
  uint64_t comp_off = 0;                 // where the blob sits in the device compressed buffer
  int tile_w, tile_h, channels, bps, ct;
  std::vector<Blit> blits;               // rectangles of the decoded tile to place in the byte output
  int status = 0;
};

// Everything a batch of tiles needs besides the decoder's unit plan: raw-plane copies, the blit table, what to report.
struct WsiWork {
  std::vector<PlaneFillJob> fills;             // plane mode 3 (raw little-endian samples)
  std::vector<TileBlitJob> blits;
  std::vector<std::pair<int, int>> unit_tile;  // unit -> tile
  uint64_t ptot = 0;                           // plane elements (u16) to reserve
};

// decompressTileBlob for a batch of tiles (wsicompress.go:424-484): every compressed plane becomes one unit of the
// decoder plan, raw planes a copy job, constant planes (modes 0 and 1) a field of the tile's blit jobs.
// T.comp_off must be set; nothing touches the device here.
void wsi_build(micgpu_decoder* d, std::vector<TileReq>& tiles, size_t out_bytes, WsiWork& W) {
  d->units.clear();
  d->temporal.clear();
  d->zero_ranges.clear();
  d->out_need = 0;
  for (size_t t = 0; t < tiles.size(); t++) {
    TileReq& T = tiles[t];
    if (T.status) continue;
    const bool rgb = T.channels == 3 && T.bps == 8;
    if (!rgb && T.channels != 1) { T.status = fail(MICGPU_E_UNSUPPORTED, "MIC3: %d channels at %d bits is not supported", T.channels, T.bps); continue; }
    const uint64_t px = (uint64_t)T.tile_w * T.tile_h;
    const int nplanes = rgb ? 3 : 1;
    const uint8_t* pl[3]; size_t ln[3];
    if (rgb) {
      if (T.len < 12) { T.status = fail(MICGPU_E_HEADER, "MIC3: RGB tile blob too small"); continue; }
      const size_t l0 = rd32(T.blob), l1 = rd32(T.blob + 4), l2 = rd32(T.blob + 8);
      if (l0 > T.len - 12 || l1 > T.len - 12 - l0 || l2 > T.len - 12 - l0 - l1) { T.status = fail(MICGPU_E_HEADER, "MIC3: RGB tile blob truncated"); continue; }
      pl[0] = T.blob + 12; ln[0] = l0; pl[1] = pl[0] + l0; ln[1] = l1; pl[2] = pl[1] + l1; ln[2] = l2;
    } else {
      pl[0] = T.blob; ln[0] = T.len;
    }
    const size_t units_before = d->units.size(), fills_before = W.fills.size(), ut_before = W.unit_tile.size(), blits_before = W.blits.size();
    const uint64_t ptot_before = W.ptot;
    TileBlitJob proto;
    memset(&proto, 0, sizeof proto);
    for (int k = 0; k < nplanes && !T.status; k++) {
      const uint64_t boff = T.comp_off + (uint64_t)(pl[k] - T.blob);
      if (ln[k] == 0) { T.status = fail(MICGPU_E_HEADER, "empty plane data"); break; }
      switch (pl[k][0]) {   // decompressWSIPlane (wsicompress.go:487-524)
        case 0: proto.cmask |= 1u << k; proto.cval[k] = 0; break;
        case 1:
          if (ln[k] < 3) { T.status = fail(MICGPU_E_HEADER, "constant plane data truncated"); break; }
          proto.cmask |= 1u << k; proto.cval[k] = (unsigned short)(pl[k][1] | (pl[k][2] << 8));
          break;
        case 2:
          proto.plane_off[k] = W.ptot;
          add_unit_locked(d, pl[k] + 1, ln[k] - 1, boff + 1, MIC_KIND_SPATIAL, (uint32_t)T.tile_w, (uint32_t)T.tile_h, W.ptot);
          W.unit_tile.push_back({(int)d->units.size() - 1, (int)t});
          W.ptot += (px + 7) & ~7ull;     // planes start on 16 B boundaries (vector blit)
          break;
        case 3:
          if (ln[k] < 1 + px * 2) { T.status = fail(MICGPU_E_HEADER, "raw plane data truncated"); break; }
          proto.plane_off[k] = W.ptot;
          W.fills.push_back(PlaneFillJob{W.ptot, boff + 1, px, 1, 0});
          W.ptot += (px + 7) & ~7ull;
          break;
        default: T.status = fail(MICGPU_E_HEADER, "unknown plane mode %d", pl[k][0]);
      }
    }
    const unsigned bppx = rgb ? 3u : (T.bps <= 8 ? 1u : 2u);
    if (!T.status) {
      proto.tile_w = (unsigned)T.tile_w; proto.tile_h = (unsigned)T.tile_h;
      proto.mode = rgb ? (T.ct ? 0u : 1u) : (T.bps <= 8 ? 2u : 3u);
      for (const Blit& b : T.blits) {
        // a rectangle must lie inside the tile and inside the byte output (k_tile_blit does not clip)
        if ((uint64_t)b.sx + b.w > (uint64_t)T.tile_w || (uint64_t)b.sy + b.h > (uint64_t)T.tile_h || b.w == 0 || b.h == 0 ||
            (uint64_t)b.w * bppx > b.dst_pitch || b.dst_off + (uint64_t)(b.h - 1) * b.dst_pitch + (uint64_t)b.w * bppx > out_bytes) {
          T.status = fail(MICGPU_E_SIZE, "MIC3: blit rectangle outside the tile or the output");
          break;
        }
        TileBlitJob J = proto;
        J.dst_off = b.dst_off; J.src_x = b.sx; J.src_y = b.sy; J.copy_w = b.w; J.copy_h = b.h; J.dst_pitch = b.dst_pitch;
        W.blits.push_back(J);
      }
    }
    if (T.status) {   // a tile that failed planning leaves nothing behind
      d->units.resize(units_before);
      W.fills.resize(fills_before);
      W.unit_tile.resize(ut_before);
      W.blits.resize(blits_before);
      W.ptot = ptot_before;
    }
  }
}

// Upload the job tables (once per plan) behind the decoder's own buffers.
int wsi_upload_jobs(micgpu_decoder* d, const WsiWork& W, cudaStream_t st) {
  int rc;
  const size_t fbytes = W.fills.size() * sizeof(PlaneFillJob), bbytes = W.blits.size() * sizeof(TileBlitJob);
  if ((rc = d->d_jobs.ensure(((fbytes + 15) & ~(size_t)15) + bbytes + 64))) return rc;
  if (fbytes) CUDA_TRY(cudaMemcpyAsync(d->d_jobs.p, W.fills.data(), fbytes, cudaMemcpyHostToDevice, st));
  if (bbytes) CUDA_TRY(cudaMemcpyAsync((uint8_t*)d->d_jobs.p + ((fbytes + 15) & ~(size_t)15), W.blits.data(), bbytes, cudaMemcpyHostToDevice, st));
  return 0;
}

// K1..K4 over the planned units, raw-plane copies, then the blits into d_bytes_out.
int wsi_launch(micgpu_decoder* d, const WsiWork& W, const void* d_comp, size_t comp_bytes, void* d_planes, void* d_bytes_out, cudaStream_t st) {
  int rc;
  if ((rc = run_device_locked(d, d_comp, comp_bytes, d_planes, (size_t)W.ptot, st))) return rc;
  const size_t fbytes = W.fills.size() * sizeof(PlaneFillJob);
  launch_plane_fill((const PlaneFillJob*)d->d_jobs.p, (int)W.fills.size(), (const uint8_t*)d_comp, (uint16_t*)d_planes, st);
  launch_tile_blit((const TileBlitJob*)((uint8_t*)d->d_jobs.p + ((fbytes + 15) & ~(size_t)15)), (int)W.blits.size(), (const uint16_t*)d_planes,
                   (uint8_t*)d_bytes_out, st);
  d->launches += (W.fills.empty() ? 0 : 1) + (W.blits.empty() ? 0 : 1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// One-shot path of the tile / region / RGB calls: blobs are staged one by one (they may be scattered in the container),
// output bytes land in d->d_bytes; the caller copies them back.
int wsi_run_locked(micgpu_decoder* d, std::vector<TileReq>& tiles, size_t out_bytes) {
  uint64_t ctot = 0;
  for (TileReq& T : tiles) {
    T.comp_off = ctot;
    ctot += (T.len + 63) & ~(size_t)63;
  }
  WsiWork W;
  wsi_build(d, tiles, out_bytes, W);
  int rc = plan_commit(d);
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(d->device));
  if ((rc = d->d_comp.ensure(ctot + 256))) return rc;
  if ((rc = d->d_out.ensure(std::max<uint64_t>(W.ptot, 1) * sizeof(uint16_t)))) return rc;
  if ((rc = d->d_bytes.ensure(std::max<size_t>(out_bytes, 1)))) return rc;
  cudaStream_t st = d->stream;
  for (size_t t = 0; t < tiles.size(); t++)
    if (!tiles[t].status) CUDA_TRY(cudaMemcpyAsync((uint8_t*)d->d_comp.p + tiles[t].comp_off, tiles[t].blob, tiles[t].len, cudaMemcpyHostToDevice, st));
  if ((rc = wsi_upload_jobs(d, W, st))) return rc;
  if ((rc = wsi_launch(d, W, d->d_comp.p, ctot, d->d_out.p, d->d_bytes.p, st))) return rc;
  std::vector<int> ust(d->units.size());
  unit_status_locked(d, ust.data(), (int)ust.size(), st);
  for (auto& ut : W.unit_tile)
    if (ust[ut.first] && !tiles[ut.second].status) tiles[ut.second].status = ust[ut.first];
  return 0;
}

int bytes_per_pixel(int channels, int bps) { return bps == 16 ? channels * 2 : channels; }

}  // namespace

extern "C" {

int micgpu_wsi_read_header(const uint8_t* mic3, size_t len, micgpu_wsi_info* info) {
  if (!mic3 || !info) return fail(MICGPU_E_HEADER, "null argument");
  Mic3Header h;
  int rc = parse_mic3(mic3, len, h);
  if (rc) return rc;
  memset(info, 0, sizeof *info);
  info->width = h.width; info->height = h.height; info->tile_w = h.tile_w; info->tile_h = h.tile_h;
  info->channels = h.channels; info->bits_per_sample = h.bps; info->color_transform = h.ct; info->n_levels = h.nlv;
  info->total_tiles = h.total_tiles;
  for (int l = 0; l < h.nlv && l < MICGPU_WSI_MAX_LEVELS; l++) {
    const Mic3Level lv = mic3_level(mic3, h, l);
    info->level_w[l] = lv.w; info->level_h[l] = lv.h; info->tiles_x[l] = lv.tx; info->tiles_y[l] = lv.ty; info->first_tile[l] = lv.first;
  }
  return 0;
}

int micgpu_wsi_decompress_tiles(const uint8_t* mic3, size_t len, int n, const int* levels, const int* txs, const int* tys,
                                uint8_t* const* outs, const size_t* caps, int* widths, int* heights, int* status) {
  if (!mic3 || n < 0) return fail(MICGPU_E_HEADER, "bad argument");
  if (n == 0) return 0;
  Mic3Header h;
  int rc = parse_mic3(mic3, len, h);
  if (rc) return rc;
  micgpu_decoder* d = default_decoder(current_device());
  if (!d) return MICGPU_E_CUDA;
  std::lock_guard<std::mutex> lk(d->mu);
  const int bpp = bytes_per_pixel(h.channels, h.bps);
  std::vector<TileReq> tiles(n);
  std::vector<size_t> ooff(n), osz(n);
  size_t otot = 0;
  for (int i = 0; i < n; i++) {
    TileReq& T = tiles[i];
    T.tile_w = h.tile_w; T.tile_h = h.tile_h; T.channels = h.channels; T.bps = h.bps; T.ct = h.ct;
    T.blob = mic3; T.len = 0;
    ooff[i] = otot; osz[i] = 0;
    if (levels[i] < 0 || levels[i] >= h.nlv) { T.status = fail(MICGPU_E_HEADER, "MIC3: level %d out of range [0, %d)", levels[i], h.nlv); continue; }
    const Mic3Level lv = mic3_level(mic3, h, levels[i]);
    if (txs[i] < 0 || txs[i] >= lv.tx || tys[i] < 0 || tys[i] >= lv.ty) { T.status = fail(MICGPU_E_HEADER, "MIC3: tile (%d,%d) out of range for level %d", txs[i], tys[i], levels[i]); continue; }
    if ((T.status = mic3_tile_blob(mic3, len, h, (long long)lv.first + (long long)tys[i] * lv.tx + txs[i], &T.blob, &T.len))) continue;
    // edge tiles are cropped to the level extent (wsicompress.go:200-216)
    const int aw = (int)std::min<long long>(h.tile_w, (long long)lv.w - (long long)txs[i] * h.tile_w);
    const int ah = (int)std::min<long long>(h.tile_h, (long long)lv.h - (long long)tys[i] * h.tile_h);
    if (aw <= 0 || ah <= 0) { T.status = fail(MICGPU_E_HEADER, "MIC3: tile (%d,%d) lies outside level %d", txs[i], tys[i], levels[i]); continue; }
    if (widths) widths[i] = aw;
    if (heights) heights[i] = ah;
    osz[i] = (size_t)aw * ah * bpp;
    if (osz[i] > caps[i]) { T.status = fail(MICGPU_E_SIZE, "tile %d: output buffer too small", i); osz[i] = 0; continue; }
    T.blits.push_back(Blit{0, 0, (unsigned)aw, (unsigned)ah, otot, (unsigned)(aw * bpp)});
    otot += (osz[i] + 15) & ~(size_t)15;
  }
  if ((rc = wsi_run_locked(d, tiles, otot))) return rc;
  int first = 0;
  for (int i = 0; i < n; i++) {
    if (!tiles[i].status && osz[i]) CUDA_TRY(cudaMemcpyAsync(outs[i], (uint8_t*)d->d_bytes.p + ooff[i], osz[i], cudaMemcpyDeviceToHost, d->stream));
    if (status) status[i] = tiles[i].status;
    if (!first && tiles[i].status) first = tiles[i].status;
  }
  CUDA_TRY(cudaStreamSynchronize(d->stream));
  return first;
}

int micgpu_wsi_decompress_tile(const uint8_t* mic3, size_t len, int level, int tx, int ty, uint8_t* out, size_t cap, int* width, int* height) {
  int w = 0, h = 0, st = 0;
  const int rc = micgpu_wsi_decompress_tiles(mic3, len, 1, &level, &tx, &ty, &out, &cap, &w, &h, &st);
  if (width) *width = w;
  if (height) *height = h;
  return rc;
}

// DecompressWSIRegion (wsicompress.go:220-296)
int micgpu_wsi_decompress_region(const uint8_t* mic3, size_t len, int level, int x, int y, int w, int h, uint8_t* out, size_t cap,
                                 int* out_w, int* out_h) {
  if (!mic3 || !out) return fail(MICGPU_E_HEADER, "null argument");
  Mic3Header hd;
  int rc = parse_mic3(mic3, len, hd);
  if (rc) return rc;
  if (level < 0 || level >= hd.nlv) return fail(MICGPU_E_HEADER, "MIC3: level %d out of range [0, %d)", level, hd.nlv);
  const Mic3Level lv = mic3_level(mic3, hd, level);
  if (x < 0 || y < 0) return fail(MICGPU_E_HEADER, "MIC3: negative region origin");
  if ((long long)x + w > lv.w) w = (int)std::max<long long>(0, (long long)lv.w - x);
  if ((long long)y + h > lv.h) h = (int)std::max<long long>(0, (long long)lv.h - y);
  if (w <= 0 || h <= 0) return fail(MICGPU_E_HEADER, "MIC3: empty region");
  const int bpp = bytes_per_pixel(hd.channels, hd.bps);
  if ((size_t)w * h * bpp > cap) return fail(MICGPU_E_SIZE, "region output buffer too small");
  if (out_w) *out_w = w;
  if (out_h) *out_h = h;
  micgpu_decoder* d = default_decoder(current_device());
  if (!d) return MICGPU_E_CUDA;
  std::lock_guard<std::mutex> lk(d->mu);
  const int stx = x / hd.tile_w, sty = y / hd.tile_h, etx = (x + w - 1) / hd.tile_w, ety = (y + h - 1) / hd.tile_h;
  std::vector<TileReq> tiles;
  for (int ty = sty; ty <= ety; ty++)
    for (int tx = stx; tx <= etx; tx++) {
      TileReq T;
      T.tile_w = hd.tile_w; T.tile_h = hd.tile_h; T.channels = hd.channels; T.bps = hd.bps; T.ct = hd.ct;
      if ((rc = mic3_tile_blob(mic3, len, hd, (long long)lv.first + (long long)ty * lv.tx + tx, &T.blob, &T.len))) return rc;
      const int tsx = tx * hd.tile_w, tsy = ty * hd.tile_h;
      const int tw = std::min(hd.tile_w, lv.w - tsx), th = std::min(hd.tile_h, lv.h - tsy);
      const int ox0 = std::max(x, tsx), oy0 = std::max(y, tsy), ox1 = std::min(x + w, tsx + tw), oy1 = std::min(y + h, tsy + th);
      if (ox1 > ox0 && oy1 > oy0)
        T.blits.push_back(Blit{(unsigned)(ox0 - tsx), (unsigned)(oy0 - tsy), (unsigned)(ox1 - ox0), (unsigned)(oy1 - oy0),
                               ((unsigned long long)(oy0 - y) * w + (ox0 - x)) * bpp, (unsigned)(w * bpp)});
      tiles.push_back(std::move(T));
    }
  const size_t obytes = (size_t)w * h * bpp;
  if ((rc = wsi_run_locked(d, tiles, obytes))) return rc;
  for (auto& T : tiles)
    if (T.status) return T.status;
  CUDA_TRY(cudaMemcpyAsync(out, d->d_bytes.p, obytes, cudaMemcpyDeviceToHost, d->stream));
  CUDA_TRY(cudaStreamSynchronize(d->stream));
  return 0;
}

// DecompressRGB (rgbcompress.go:31-33): one whole-image "tile" with the colour transform on
int micgpu_rgb_decompress(const uint8_t* blob, size_t len, int width, int height, uint8_t* rgb_out) {
  if (!blob || !rgb_out || width <= 0 || height <= 0) return fail(MICGPU_E_HEADER, "bad argument");
  micgpu_decoder* d = default_decoder(current_device());
  if (!d) return MICGPU_E_CUDA;
  std::lock_guard<std::mutex> lk(d->mu);
  std::vector<TileReq> tiles(1);
  TileReq& T = tiles[0];
  T.blob = blob; T.len = len; T.tile_w = width; T.tile_h = height; T.channels = 3; T.bps = 8; T.ct = 1;
  T.blits.push_back(Blit{0, 0, (unsigned)width, (unsigned)height, 0, (unsigned)(width * 3)});
  const size_t obytes = (size_t)width * height * 3;
  int rc = wsi_run_locked(d, tiles, obytes);
  if (rc) return rc;
  if (T.status) return T.status;
  CUDA_TRY(cudaMemcpyAsync(rgb_out, d->d_bytes.p, obytes, cudaMemcpyDeviceToHost, d->stream));
  CUDA_TRY(cudaStreamSynchronize(d->stream));
  return 0;
}

}  // extern "C"

extern "C" {
}  // extern "C"

// ---- MIC3 tile ranges: plan once, run many (batch / serving path) ---------------------------------------------------
struct micgpu_wsi_plan {
  micgpu_decoder* dec = nullptr;    // owns the unit plan, the scratch and the job tables
  bool own_dec = false;
  WsiWork W;
  uint64_t first = 0, n = 0;
  uint64_t span_off = 0, span_len = 0;     // bytes of the container that hold the range's tile blobs
  uint64_t tile_bytes = 0, out_bytes = 0;
  std::vector<int> host_status;            // per tile: what planning found
};

namespace {

int wsi_plan_build(micgpu_wsi_plan* P, const uint8_t* mic3, size_t len, uint64_t first, uint64_t n) {
  Mic3Header h;
  int rc = parse_mic3(mic3, len, h);
  if (rc) return rc;
  if (first > h.total_tiles || n > h.total_tiles - first) return fail(MICGPU_E_HEADER, "MIC3: tile range %llu+%llu exceeds %llu tiles", (unsigned long long)first, (unsigned long long)n, (unsigned long long)h.total_tiles);
  if (n > 0x7FFFFFFFull / 4) return fail(MICGPU_E_UNSUPPORTED, "MIC3: %llu tiles in one plan", (unsigned long long)n);
  const int bpp = h.bps == 16 ? h.channels * 2 : h.channels;
  P->first = first; P->n = n;
  P->tile_bytes = (uint64_t)h.tile_w * h.tile_h * bpp;
  P->out_bytes = P->tile_bytes * n;
  P->host_status.assign(n, 0);
  std::vector<TileReq> tiles(n);
  uint64_t lo = ~0ull, hi = 0;
  for (uint64_t t = 0; t < n; t++) {
    TileReq& T = tiles[t];
    T.tile_w = h.tile_w; T.tile_h = h.tile_h; T.channels = h.channels; T.bps = h.bps; T.ct = h.ct;
    T.blob = mic3; T.len = 0;
    if ((T.status = mic3_tile_blob(mic3, len, h, (long long)(first + t), &T.blob, &T.len))) continue;
    const uint64_t o = (uint64_t)(T.blob - mic3);
    lo = std::min(lo, o);
    hi = std::max(hi, o + T.len);
    T.blits.push_back(Blit{0, 0, (unsigned)h.tile_w, (unsigned)h.tile_h, t * P->tile_bytes, (unsigned)(h.tile_w * bpp)});
  }
  if (hi <= lo) { lo = 0; hi = 0; }
  P->span_off = lo; P->span_len = hi - lo;
  for (TileReq& T : tiles)
    if (!T.status) T.comp_off = (uint64_t)(T.blob - mic3) - lo;
  micgpu_decoder* d = P->dec;
  P->W = WsiWork();
  wsi_build(d, tiles, (size_t)P->out_bytes, P->W);
  for (uint64_t t = 0; t < n; t++) P->host_status[t] = tiles[t].status;
  if ((rc = plan_commit(d))) return rc;
  CUDA_TRY(cudaSetDevice(d->device));
  if ((rc = d->d_out.ensure(std::max<uint64_t>(P->W.ptot, 1) * sizeof(uint16_t)))) return rc;
  if ((rc = wsi_upload_jobs(d, P->W, d->stream))) return rc;
  CUDA_TRY(cudaStreamSynchronize(d->stream));
  return 0;
}

int wsi_plan_fold_status(micgpu_wsi_plan* P, int* tile_status, int n, cudaStream_t st) {
  micgpu_decoder* d = P->dec;
  std::vector<int> ust(d->units.size());
  unit_status_locked(d, ust.data(), (int)ust.size(), st);
  std::vector<int> ts(P->host_status);
  for (auto& ut : P->W.unit_tile)
    if (ust[ut.first] && !ts[ut.second]) ts[ut.second] = ust[ut.first];
  int first = 0;
  for (size_t t = 0; t < ts.size(); t++) {
    if (tile_status && (int)t < n) tile_status[t] = ts[t];
    if (!first && ts[t]) first = ts[t];
  }
  return first;
}

}  // namespace

extern "C" {

micgpu_wsi_plan* micgpu_wsi_plan_tiles(int device, const uint8_t* mic3, size_t len, uint64_t first_tile, uint64_t n_tiles) {
  if (!mic3) { fail(MICGPU_E_HEADER, "null argument"); return nullptr; }
  micgpu_decoder* d = micgpu_decoder_create(device);
  if (!d) return nullptr;
  micgpu_wsi_plan* P = new micgpu_wsi_plan();
  P->dec = d;
  P->own_dec = true;
  std::lock_guard<std::mutex> lk(d->mu);
  if (wsi_plan_build(P, mic3, len, first_tile, n_tiles)) {
    delete d;
    delete P;
    return nullptr;
  }
  return P;
}

void micgpu_wsi_plan_destroy(micgpu_wsi_plan* p) {
  if (!p) return;
  if (p->own_dec) delete p->dec;
  delete p;
}

int micgpu_wsi_plan_info(const micgpu_wsi_plan* p, uint64_t* span_off, uint64_t* span_len, uint64_t* tile_bytes, uint64_t* out_bytes,
                         int* n_units) {
  if (!p) return fail(MICGPU_E_HEADER, "null plan");
  if (span_off) *span_off = p->span_off;
  if (span_len) *span_len = p->span_len;
  if (tile_bytes) *tile_bytes = p->tile_bytes;
  if (out_bytes) *out_bytes = p->out_bytes;
  if (n_units) *n_units = (int)p->dec->units.size();
  return 0;
}

int micgpu_wsi_plan_run_device(micgpu_wsi_plan* p, const void* d_span, void* d_out, void* cuda_stream) {
  if (!p || (!d_span && p->span_len) || (!d_out && p->out_bytes)) return fail(MICGPU_E_HEADER, "null argument");
  micgpu_decoder* d = p->dec;
  std::lock_guard<std::mutex> lk(d->mu);
  CUDA_TRY(cudaSetDevice(d->device));
  return wsi_launch(d, p->W, d_span, (size_t)p->span_len, d->d_out.p, d_out, (cudaStream_t)cuda_stream);
}

int micgpu_wsi_plan_status(micgpu_wsi_plan* p, int* tile_status, int n, void* cuda_stream) {
  if (!p) return fail(MICGPU_E_HEADER, "null plan");
  std::lock_guard<std::mutex> lk(p->dec->mu);
  return wsi_plan_fold_status(p, tile_status, n, (cudaStream_t)cuda_stream);
}

int micgpu_wsi_plan_launches(const micgpu_wsi_plan* p) { return p ? p->dec->launches : 0; }

int micgpu_wsi_plan_kernel_times(micgpu_wsi_plan* p, int on, char* names, size_t names_cap, float* ms, int cap) {
  if (!p) return fail(MICGPU_E_HEADER, "null plan");
  if (on >= 0) return micgpu_decoder_set_profiling(p->dec, on);
  return micgpu_decoder_kernel_times(p->dec, names, names_cap, ms, cap);
}

}  // extern "C"

// ---- MIC3 region / viewport serving: the slide stays resident, headers are parsed once ----------------
// DecompressWSIRegion (wsicompress.go:220-296) re-reads the header and re-extracts blobs on every call; a viewer asks for
// many small rectangles of one slide, often from several pyramid levels at once.  A slide handle validates the header,
// the level descriptors and the tile table ONCE, keeps the tile data area in device memory (a 100000 x 80000 slide is a
// few GB of the 180 GB), and serves a batch of rectangles -- any levels -- as one launch sequence: the union of the
// tiles they touch is decoded once (a tile shared by two rectangles is one unit set with two blit jobs), cropped and
// composed on the device, and only the requested pixels travel back.
struct micgpu_wsi_slide {
  micgpu_decoder* dec = nullptr;
  std::vector<uint8_t> file;       // private host copy: tile blob headers (plane lengths, modes, FSE framing) are read per request
  Mic3Header h;
  std::vector<Mic3Level> levels;
  DevBuf d_data;                   // file[data_off .. len) (+256 readable bytes)
  uint64_t requests = 0, tiles_decoded = 0;
  ~micgpu_wsi_slide() {
    if (dec) { cudaSetDevice(dec->device); d_data.release(); delete dec; }
  }
};

extern "C" {

micgpu_wsi_slide* micgpu_wsi_open(int device, const uint8_t* mic3, size_t len) {
  if (!mic3) { fail(MICGPU_E_HEADER, "null argument"); return nullptr; }
  Mic3Header h;
  if (parse_mic3(mic3, len, h)) return nullptr;
  micgpu_decoder* d = micgpu_decoder_create(device);
  if (!d) return nullptr;
  micgpu_wsi_slide* S = new micgpu_wsi_slide();
  S->dec = d;
  S->h = h;
  S->file.assign(mic3, mic3 + len);
  for (int l = 0; l < h.nlv; l++) S->levels.push_back(mic3_level(mic3, h, l));
  const size_t dbytes = len - h.data_off;
  if (S->d_data.ensure(dbytes + 256) ||
      cudaMemcpyAsync(S->d_data.p, mic3 + h.data_off, dbytes, cudaMemcpyHostToDevice, d->stream) != cudaSuccess ||
      cudaMemsetAsync((uint8_t*)S->d_data.p + dbytes, 0, 256, d->stream) != cudaSuccess || cudaStreamSynchronize(d->stream) != cudaSuccess) {
    fail(MICGPU_E_CUDA, "MIC3: could not place %zu bytes of tile data in device memory", dbytes);
    delete S;
    return nullptr;
  }
  return S;
}

void micgpu_wsi_close(micgpu_wsi_slide* s) { delete s; }

int micgpu_wsi_slide_info(const micgpu_wsi_slide* s, micgpu_wsi_info* info) {
  if (!s || !info) return fail(MICGPU_E_HEADER, "null argument");
  return micgpu_wsi_read_header(s->file.data(), s->file.size(), info);
}

// n rectangles (level, x, y, w, h), clamped to their level like DecompressWSIRegion; outs[i] receives out_ws[i] x
// out_hs[i] pixels, row-major, bytes_per_pixel each.  status[i] (optional) is the per-rectangle result; the return value
// is the first failure.  d_outs != NULL: outs[] are DEVICE pointers (a renderer that keeps pixels on the GPU).
static int slide_regions(micgpu_wsi_slide* S, int n, const int* levels, const int* xs, const int* ys, const int* ws, const int* hs,
                         uint8_t* const* outs, const size_t* caps, int* out_ws, int* out_hs, int* status, bool device_out) {
  if (!S || n < 0 || (n && (!levels || !xs || !ys || !ws || !hs || !outs || !caps))) return fail(MICGPU_E_HEADER, "bad argument");
  if (n == 0) return 0;
  micgpu_decoder* d = S->dec;
  std::lock_guard<std::mutex> lk(d->mu);
  const Mic3Header& hd = S->h;
  const uint8_t* file = S->file.data();
  const size_t flen = S->file.size();
  const int bpp = bytes_per_pixel(hd.channels, hd.bps);
  std::vector<TileReq> tiles;
  std::vector<std::vector<int>> region_tiles(n);
  std::unordered_map<uint64_t, int> tile_of;       // tile table index -> entry of `tiles`
  std::vector<int> rst(n, 0), rw(n, 0), rh(n, 0);
  std::vector<size_t> ooff(n, 0);
  size_t otot = 0;
  for (int i = 0; i < n; i++) {
    const int level = levels[i], x = xs[i], y = ys[i];
    int w = ws[i], h = hs[i];
    if (level < 0 || level >= hd.nlv) { rst[i] = fail(MICGPU_E_HEADER, "MIC3: level %d out of range [0, %d)", level, hd.nlv); continue; }
    const Mic3Level& lv = S->levels[level];
    if (x < 0 || y < 0) { rst[i] = fail(MICGPU_E_HEADER, "MIC3: negative region origin"); continue; }
    if ((long long)x + w > lv.w) w = (int)std::max<long long>(0, (long long)lv.w - x);
    if ((long long)y + h > lv.h) h = (int)std::max<long long>(0, (long long)lv.h - y);
    if (w <= 0 || h <= 0) { rst[i] = fail(MICGPU_E_HEADER, "MIC3: empty region"); continue; }
    if ((size_t)w * h * bpp > caps[i]) { rst[i] = fail(MICGPU_E_SIZE, "region %d: output buffer too small", i); continue; }
    rw[i] = w; rh[i] = h;
    ooff[i] = otot;
    const int stx = x / hd.tile_w, sty = y / hd.tile_h, etx = (x + w - 1) / hd.tile_w, ety = (y + h - 1) / hd.tile_h;
    for (int ty = sty; ty <= ety && !rst[i]; ty++)
      for (int tx = stx; tx <= etx; tx++) {
        const uint64_t idx = (uint64_t)lv.first + (uint64_t)ty * lv.tx + tx;
        auto it = tile_of.find(idx);
        int ti;
        if (it == tile_of.end()) {
          TileReq T;
          T.tile_w = hd.tile_w; T.tile_h = hd.tile_h; T.channels = hd.channels; T.bps = hd.bps; T.ct = hd.ct;
          T.blob = file; T.len = 0;
          if ((T.status = mic3_tile_blob(file, flen, hd, (long long)idx, &T.blob, &T.len)) == 0)
            T.comp_off = (uint64_t)(T.blob - file) - hd.data_off;
          ti = (int)tiles.size();
          tiles.push_back(std::move(T));
          tile_of.emplace(idx, ti);
        } else {
          ti = it->second;
        }
        region_tiles[i].push_back(ti);
        const int tsx = tx * hd.tile_w, tsy = ty * hd.tile_h;
        const int tw = std::min(hd.tile_w, lv.w - tsx), th = std::min(hd.tile_h, lv.h - tsy);
        const int ox0 = std::max(x, tsx), oy0 = std::max(y, tsy), ox1 = std::min(x + w, tsx + tw), oy1 = std::min(y + h, tsy + th);
        if (ox1 > ox0 && oy1 > oy0)
          tiles[ti].blits.push_back(Blit{(unsigned)(ox0 - tsx), (unsigned)(oy0 - tsy), (unsigned)(ox1 - ox0), (unsigned)(oy1 - oy0),
                                         otot + ((unsigned long long)(oy0 - y) * w + (ox0 - x)) * bpp, (unsigned)(w * bpp)});
      }
    otot += ((size_t)w * h * bpp + 15) & ~(size_t)15;
  }
  int rc = 0;
  if (!tiles.empty()) {
    WsiWork W;
    wsi_build(d, tiles, otot, W);
    if ((rc = plan_commit(d))) return rc;
    CUDA_TRY(cudaSetDevice(d->device));
    if ((rc = d->d_out.ensure(std::max<uint64_t>(W.ptot, 1) * sizeof(uint16_t)))) return rc;
    if ((rc = d->d_bytes.ensure(std::max<size_t>(otot, 1)))) return rc;
    cudaStream_t st = d->stream;
    if ((rc = wsi_upload_jobs(d, W, st))) return rc;
    if ((rc = wsi_launch(d, W, S->d_data.p, flen - hd.data_off, d->d_out.p, d->d_bytes.p, st))) { cudaStreamSynchronize(st); return rc; }
    std::vector<int> ust(d->units.size());
    unit_status_locked(d, ust.data(), (int)ust.size(), st);    // synchronises
    for (auto& ut : W.unit_tile)
      if (ust[ut.first] && !tiles[ut.second].status) tiles[ut.second].status = ust[ut.first];
    S->tiles_decoded += tiles.size();
  }
  S->requests += (uint64_t)n;
  int first = 0;
  for (int i = 0; i < n; i++) {
    for (int ti : region_tiles[i])
      if (!rst[i] && tiles[ti].status) rst[i] = tiles[ti].status;
    if (!rst[i]) {
      const size_t nb = (size_t)rw[i] * rh[i] * bpp;
      if (cudaMemcpyAsync(outs[i], (uint8_t*)d->d_bytes.p + ooff[i], nb, device_out ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, d->stream) != cudaSuccess)
        rst[i] = fail(MICGPU_E_CUDA, "copy of region %d failed", i);
    }
    if (out_ws) out_ws[i] = rst[i] ? 0 : rw[i];
    if (out_hs) out_hs[i] = rst[i] ? 0 : rh[i];
    if (status) status[i] = rst[i];
    if (!first && rst[i]) first = rst[i];
  }
  CUDA_TRY(cudaStreamSynchronize(d->stream));
  if (first) fail(first, "MIC3: a region of the batch failed (status %d)", first);
  return first;
}

int micgpu_wsi_slide_regions(micgpu_wsi_slide* s, int n, const int* levels, const int* xs, const int* ys, const int* ws, const int* hs,
                             uint8_t* const* outs, const size_t* caps, int* out_ws, int* out_hs, int* status) {
  return slide_regions(s, n, levels, xs, ys, ws, hs, outs, caps, out_ws, out_hs, status, false);
}

int micgpu_wsi_slide_regions_device(micgpu_wsi_slide* s, int n, const int* levels, const int* xs, const int* ys, const int* ws,
                                    const int* hs, void* const* d_outs, const size_t* caps, int* out_ws, int* out_hs, int* status) {
  return slide_regions(s, n, levels, xs, ys, ws, hs, (uint8_t* const*)d_outs, caps, out_ws, out_hs, status, true);
}

int micgpu_wsi_slide_region(micgpu_wsi_slide* s, int level, int x, int y, int w, int h, uint8_t* out, size_t cap, int* out_w, int* out_h) {
  int st = 0;
  return slide_regions(s, 1, &level, &x, &y, &w, &h, &out, &cap, out_w, out_h, &st, false);
}

// counters of a handle: rectangles served, tiles decoded (after sharing), kernels of the last batch
int micgpu_wsi_slide_stats(const micgpu_wsi_slide* s, uint64_t* requests, uint64_t* tiles_decoded, int* last_launches) {
  if (!s) return fail(MICGPU_E_HEADER, "null slide");
  if (requests) *requests = s->requests;
  if (tiles_decoded) *tiles_decoded = s->tiles_decoded;
  if (last_launches) *last_launches = s->dec->launches;
  return 0;
}

}  // extern "C"

namespace {

// One device's share of a tile-range call: plan, one H2D copy of the span, kernels, one D2H copy.
int wsi_range_on_device(int dev, int slot, const uint8_t* mic3, size_t len, uint64_t first, uint64_t n, uint8_t* out, int* status, std::string* msg) {
  auto done = [&](int rc) { if (rc && msg) *msg = err_slot(); return rc; };
  if (n == 0) return 0;
  if (cudaSetDevice(dev) != cudaSuccess) return done(fail(MICGPU_E_CUDA, "cudaSetDevice(%d) failed", dev));
  micgpu_decoder* d = multi_decoder(dev, 32 + slot);   // contexts of their own: the plan keeps its scratch between calls
  if (!d) return done(MICGPU_E_CUDA);
  std::lock_guard<std::mutex> lk(d->mu);
  micgpu_wsi_plan P;
  P.dec = d;
  int rc = wsi_plan_build(&P, mic3, len, first, n);
  if (rc) return done(rc);
  if ((rc = d->d_comp.ensure(P.span_len + 256))) return done(rc);
  if ((rc = d->d_bytes.ensure(std::max<uint64_t>(P.out_bytes, 1)))) return done(rc);
  cudaStream_t st = d->stream;
  if (P.span_len && cudaMemcpyAsync(d->d_comp.p, mic3 + P.span_off, P.span_len, cudaMemcpyHostToDevice, st) != cudaSuccess)
    return done(fail(MICGPU_E_CUDA, "H2D copy of the tile span failed"));
  if ((rc = wsi_launch(d, P.W, d->d_comp.p, (size_t)P.span_len, d->d_out.p, d->d_bytes.p, st))) { cudaStreamSynchronize(st); return done(rc); }
  if (cudaMemcpyAsync(out, d->d_bytes.p, P.out_bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess) {
    cudaStreamSynchronize(st);
    return done(fail(MICGPU_E_CUDA, "D2H copy of the tiles failed"));
  }
  rc = wsi_plan_fold_status(&P, status, (int)n, st);   // synchronises the stream
  return done(rc);
}

}  // namespace

extern "C" {

// Tiles [first_tile, first_tile + n_tiles) of the container's tile table, each as a full tile_w x tile_h block
// (tile_bytes apart, edge tiles keep the zero padding the encoder added, wsicompress.go:529-556).  With several
// devices configured (micgpu_init) the range is cut by compressed bytes from the u64 offset table
// (wsiformat.go:145-155) and every device decodes its share on its own host thread.
int micgpu_partition_by_bytes(const uint64_t* sizes, uint64_t n, int parts, uint64_t* cuts) {
  if ((!sizes && n) || !cuts || parts <= 0) return fail(MICGPU_E_HEADER, "bad argument");
  std::vector<uint64_t> c;
  partition_by_bytes(sizes, n, parts, c);
  for (int i = 0; i <= parts; i++) cuts[i] = c[i];
  return 0;
}

int micgpu_wsi_decompress_tile_range(const uint8_t* mic3, size_t len, uint64_t first_tile, uint64_t n_tiles, uint8_t* out, size_t cap,
                                     int* status) {
  if (!mic3 || (!out && n_tiles)) return fail(MICGPU_E_HEADER, "null argument");
  Mic3Header h;
  int rc = parse_mic3(mic3, len, h);
  if (rc) return rc;
  if (first_tile > h.total_tiles || n_tiles > h.total_tiles - first_tile) return fail(MICGPU_E_HEADER, "MIC3: tile range exceeds the tile table");
  const uint64_t tile_bytes = (uint64_t)h.tile_w * h.tile_h * (h.bps == 16 ? h.channels * 2 : h.channels);
  if (n_tiles && tile_bytes > cap / n_tiles) return fail(MICGPU_E_SIZE, "output buffer too small for %llu tiles", (unsigned long long)n_tiles);
  std::vector<int> devs = configured_devices();
  // One device and a large range: four contexts of that device take a quarter each on their own host threads, so the
  // planning of one quarter (host work, ~3 us per tile), the H2D copy of another, the kernels of a third and the D2H copy
  // of the fourth overlap (a device list may name a device several times; MICGPU_WSI_SPLIT=1 keeps one context).
  static const int split_cfg = [] { const char* e = getenv("MICGPU_WSI_SPLIT"); int v = e ? atoi(e) : 4; return v < 1 ? 1 : (v > 8 ? 8 : v); }();
  if (devs.size() <= 1 && n_tiles >= 4096 && split_cfg > 1) devs.assign((size_t)split_cfg, devs.empty() ? current_device() : devs[0]);
  if (devs.size() <= 1 || n_tiles < 2 * devs.size()) {
    std::string msg;
    return wsi_range_on_device(devs.empty() ? current_device() : devs[0], 0, mic3, len, first_tile, n_tiles, out, status, nullptr);
  }
  std::vector<uint64_t> sizes(n_tiles), cuts;
  for (uint64_t t = 0; t < n_tiles; t++) sizes[t] = rd64(mic3 + h.table_off + (size_t)(first_tile + t) * 16 + 8) + 1024;
  partition_by_bytes(sizes.data(), n_tiles, (int)devs.size(), cuts);
  std::vector<int> rcs(devs.size(), 0);
  std::vector<std::string> msgs(devs.size());
  std::vector<std::thread> th;
  for (size_t k = 0; k < devs.size(); k++)
    th.emplace_back([&, k] {
      rcs[k] = wsi_range_on_device(devs[k], (int)k, mic3, len, first_tile + cuts[k], cuts[k + 1] - cuts[k], out + cuts[k] * tile_bytes,
                                   status ? status + cuts[k] : nullptr, &msgs[k]);
    });
  for (auto& t : th) t.join();
  for (size_t k = 0; k < devs.size(); k++)
    if (rcs[k]) { err_slot() = msgs[k]; return rcs[k]; }
  return 0;
}

}  // extern "C"

// ---- WaveletV2 ---------------------------------------------------------------------
namespace {

WaveletGeom wavelet_geom(unsigned rows, unsigned cols, int levels) {
  WaveletGeom G;
  memset(&G, 0, sizeof G);
  G.rows = rows; G.cols = cols; G.levels = levels;
  unsigned nr[10], nc[10];
  nr[0] = rows; nc[0] = cols;
  for (int l = 1; l <= levels; l++) { nr[l] = (nr[l - 1] + 1) / 2; nc[l] = (nc[l - 1] + 1) / 2; }
  unsigned pos = 0;
  int n = 0;
  auto add = [&](unsigned y0, unsigned y1, unsigned x0, unsigned x1) {   // collectSubbandOrder (waveletfsecompressu16.go:202-241)
    G.seg_start[n] = pos; G.seg_y0[n] = y0; G.seg_x0[n] = x0; G.seg_w[n] = x1 > x0 ? x1 - x0 : 1;
    if (y1 > y0 && x1 > x0) pos += (y1 - y0) * (x1 - x0);
    n++;
  };
  add(0, nr[levels], 0, nc[levels]);
  for (int l = levels; l >= 1; l--) {
    add(0, nr[l], nc[l], nc[l - 1]);
    add(nr[l], nr[l - 1], 0, nc[l]);
    add(nr[l], nr[l - 1], nc[l], nc[l - 1]);
  }
  G.nseg = n;
  return G;
}

}  // namespace

extern "C" {

// WaveletV2RLEFSEDecompressU16 / WaveletV2SIMDRLEFSEDecompressU16 (waveletfsecompressu16.go:374-421,493-534) for n streams
// variant 2: WaveletV2 (11-byte header, RLE + FSE-4, subband order, Mallat levels); variant 0: WaveletFSEDecompressU16 (11-byte
// header, FSE-4 only, raster order, interleaved in-place levels); variant 1: WaveletRLEFSEDecompressU16 (15-byte header, RLE +
// FSE-4, raster order, interleaved levels) -- waveletfsecompressu16.go:124-163, 374-421, 624-669
static int wavelet_decompress_batch(int variant, int n, const uint8_t* const* blobs, const size_t* lens, uint16_t* const* outs,
                                    const size_t* caps, int* rows_out, int* cols_out, int* status);

int micgpu_wavelet_v2_decompress_batch(int n, const uint8_t* const* blobs, const size_t* lens, uint16_t* const* outs, const size_t* caps,
                                       int* rows_out, int* cols_out, int* status) {
  return wavelet_decompress_batch(2, n, blobs, lens, outs, caps, rows_out, cols_out, status);
}

// WaveletFSEDecompressU16 (with_rle = 0) / WaveletRLEFSEDecompressU16 (with_rle = 1): the V1 layouts
int micgpu_wavelet_v1_decompress(const uint8_t* blob, size_t len, int with_rle, uint16_t* pixels_out, size_t cap_px, int* rows, int* cols) {
  if (!blob || !pixels_out) return fail(MICGPU_E_HEADER, "null argument");
  return wavelet_decompress_batch(with_rle ? 1 : 0, 1, &blob, &len, &pixels_out, &cap_px, rows, cols, nullptr);
}

static int wavelet_decompress_batch(int variant, int n, const uint8_t* const* blobs, const size_t* lens, uint16_t* const* outs,
                                    const size_t* caps, int* rows_out, int* cols_out, int* status) {
  const size_t HDR = variant == 1 ? 15 : 11;
  const bool v1 = variant != 2;
  if (n <= 0) return 0;
  micgpu_decoder* d = default_decoder(current_device());
  if (!d) return MICGPU_E_CUDA;
  std::lock_guard<std::mutex> lk(d->mu);
  d->units.clear();
  d->temporal.clear();
  d->zero_ranges.clear();
  d->out_need = 0;
  struct Img { unsigned rows, cols; int levels, unit, st; uint64_t coff, px_off; };
  std::vector<Img> im(n);
  uint64_t ctot = 0, stot = 0, ptot = 0;
  for (int i = 0; i < n; i++) {
    Img& I = im[i];
    I.st = 0; I.unit = -1; I.coff = ctot; I.px_off = ptot; I.rows = I.cols = 0; I.levels = 0;
    ctot += (lens[i] + 63) & ~(size_t)63;
    if (lens[i] < HDR + 6) { I.st = fail(MICGPU_E_HEADER, "compressed data too short"); continue; }
    I.rows = rd32(blobs[i]); I.cols = rd32(blobs[i] + 4); I.levels = blobs[i][10];
    if (rows_out) rows_out[i] = (int)I.rows;
    if (cols_out) cols_out[i] = (int)I.cols;
    const uint64_t px = (uint64_t)I.rows * I.cols;
    if (I.rows == 0 || I.cols == 0 || px > (1ull << 31) || (!v1 && I.levels > 8)) { I.st = fail(MICGPU_E_HEADER, "bad wavelet header"); continue; }
    if (v1 && I.rows > 65535) { I.st = fail(MICGPU_E_UNSUPPORTED, "V1 wavelet stream with %u rows", I.rows); continue; }
    if (blobs[i][HDR] != 0xFF || blobs[i][HDR + 1] != 0x04) { I.st = fail(MICGPU_E_HEADER, "fse4state: missing magic bytes"); continue; }
    if (px > caps[i]) { I.st = fail(MICGPU_E_SIZE, "image %d: output buffer too small", i); continue; }
    // coefficient stream: one u16 per coefficient, three per escaped coefficient; the FSE symbol count bounds it
    const uint64_t count = rd32(blobs[i] + HDR + 2);
    const uint64_t cap = variant == 0 ? std::max<uint64_t>(count, 1) : std::min<uint64_t>(3 * px, std::max<uint64_t>(px, 65535ull * count));
    if (variant == 0 && count > 3 * px) { I.st = fail(MICGPU_E_SIZE, "coefficient stream longer than three words per pixel"); continue; }
    I.unit = add_unit_locked(d, blobs[i] + HDR, lens[i] - HDR, I.coff + HDR, variant == 0 ? MIC_KIND_RAW : MIC_KIND_RLE,
                             (uint32_t)std::min<uint64_t>(cap, 0xFFFFFFFFull), 1, stot);
    stot += (cap + 15) & ~15ull;
    ptot += px;
  }
  int rc = plan_commit(d);
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(d->device));
  if ((rc = d->d_comp.ensure(ctot + 256))) return rc;
  if ((rc = d->d_out.ensure(std::max<uint64_t>(stot, 1) * sizeof(uint16_t)))) return rc;
  if ((rc = d->d_bytes.ensure(std::max<uint64_t>(ptot, 1) * sizeof(uint16_t)))) return rc;
  if ((rc = d->d_wA.ensure(std::max<uint64_t>(ptot, 1) * sizeof(int32_t)))) return rc;
  if ((rc = d->d_wB.ensure(std::max<uint64_t>(ptot, 1) * sizeof(int32_t)))) return rc;
  if ((rc = d->d_wflags.ensure((size_t)n * 2 * sizeof(int) + 64))) return rc;
  cudaStream_t st = d->stream;
  for (int i = 0; i < n; i++)
    if (!im[i].st) CUDA_TRY(cudaMemcpyAsync((uint8_t*)d->d_comp.p + im[i].coff, blobs[i], lens[i], cudaMemcpyHostToDevice, st));
  if ((rc = run_device_locked(d, d->d_comp.p, ctot, d->d_out.p, stot, st))) return rc;
  // group images with the same geometry: they share one launch sequence (grid.z = image)
  std::vector<char> done(n, 0);
  std::vector<int> uoi(n);
  int* d_flags = (int*)d->d_wflags.p;
  int* d_uoi = d_flags + n;
  for (int i = 0; i < n; i++) {
    if (done[i] || im[i].st) continue;
    // images of a group must be contiguous in the plane buffers: collect consecutive equal geometries
    int j = i, cnt = 0;
    while (j < n && !im[j].st && im[j].rows == im[i].rows && im[j].cols == im[i].cols && im[j].levels == im[i].levels) {
      uoi[cnt++] = im[j].unit;
      done[j] = 1;
      j++;
    }
    WaveletGeom G;
    if (v1) {   // raster order: one segment that covers the image
      memset(&G, 0, sizeof G);
      G.rows = im[i].rows; G.cols = im[i].cols; G.levels = im[i].levels; G.nseg = 1; G.seg_w[0] = im[i].cols;
    } else {
      G = wavelet_geom(im[i].rows, im[i].cols, im[i].levels);
    }
    CUDA_TRY(cudaMemcpyAsync(d_uoi, uoi.data(), cnt * sizeof(int), cudaMemcpyHostToDevice, st));
    launch_wavelet_decode((MicUnit*)d->d_units.p, d_uoi, cnt, (const uint16_t*)d->d_out.p, d_flags, (int32_t*)d->d_wA.p + im[i].px_off,
                          (int32_t*)d->d_wB.p + im[i].px_off, (uint16_t*)d->d_bytes.p + im[i].px_off, G, st, v1 ? 1 : 0);
    d->launches += 2 + 2 * std::max(1, G.levels);
    CUDA_TRY(cudaStreamSynchronize(st));   // d_uoi / d_flags are reused by the next group
  }
  CUDA_TRY(cudaGetLastError());
  std::vector<int> ust(d->units.size());
  unit_status_locked(d, ust.data(), (int)ust.size(), st);
  int first = 0;
  for (int i = 0; i < n; i++) {
    if (!im[i].st && im[i].unit >= 0 && ust[im[i].unit]) im[i].st = ust[im[i].unit];
    if (!im[i].st)
      CUDA_TRY(cudaMemcpyAsync(outs[i], (uint16_t*)d->d_bytes.p + im[i].px_off, (size_t)im[i].rows * im[i].cols * sizeof(uint16_t), cudaMemcpyDeviceToHost, st));
    if (status) status[i] = im[i].st;
    if (!first && im[i].st) first = im[i].st;
  }
  CUDA_TRY(cudaStreamSynchronize(st));
  if (first) fail(first, "wavelet stream decode failed (status %d)", first);
  return first;
}

int micgpu_wavelet_v2_decompress(const uint8_t* blob, size_t len, uint16_t* pixels_out, size_t cap_px, int* rows, int* cols) {
  if (!blob || !pixels_out) return fail(MICGPU_E_HEADER, "null argument");
  return micgpu_wavelet_v2_decompress_batch(1, &blob, &len, &pixels_out, &cap_px, rows, cols, nullptr);
}

}  // extern "C"

extern "C" {
// ---- reference C twin symbols ----------------------------------------------------
static int twin_frame(const uint8_t* c, size_t n, uint16_t* o, int w, int h, uint8_t magic) {
  // the C twin checks its magic byte and returns -1 (ojph/mic_decompress_c.c:1004-1010)
  if (!c || n < 6 || c[0] != 0xFF || c[1] != magic) return fail(MICGPU_E_HEADER, "missing magic bytes");
  return micgpu_decompress_single_frame(c, n, o, w, h);
}
int mic_decompress_two_state(const uint8_t* c, size_t n, uint16_t* o, int w, int h) { return twin_frame(c, n, o, w, h, 0x02); }
int mic_decompress_two_state_simd(const uint8_t* c, size_t n, uint16_t* o, int w, int h) { return twin_frame(c, n, o, w, h, 0x02); }
int mic_decompress_four_state(const uint8_t* c, size_t n, uint16_t* o, int w, int h) { return twin_frame(c, n, o, w, h, 0x04); }
int mic_decompress_four_state_simd(const uint8_t* c, size_t n, uint16_t* o, int w, int h) { return twin_frame(c, n, o, w, h, 0x04); }
int mic_decompress_eight_state(const uint8_t* c, size_t n, uint16_t* o, int w, int h) { return twin_frame(c, n, o, w, h, 0x84); }
int mic_decompress_eight_state_simd(const uint8_t* c, size_t n, uint16_t* o, int w, int h) { return twin_frame(c, n, o, w, h, 0x84); }

int mic_decompress_parallel(const uint8_t* c, size_t n, uint16_t* o, int w, int h, int max_threads) {
  (void)max_threads;
  if (!c || !o) return fail(MICGPU_E_HEADER, "null argument");
  PicsHeader ph;
  int rc = parse_pics(c, n, ph);
  if (rc) return rc;
  if (ph.w != w || ph.h != h) return fail(MICGPU_E_HEADER, "header dims %dx%d != args %dx%d", ph.w, ph.h, w, h);   // mic_parallel.c:99
  int ow, oh;
  return micgpu_pics_decompress(c, n, o, (size_t)w * h, &ow, &oh);
}
int mic_decompress_parallel_scalar(const uint8_t* c, size_t n, uint16_t* o, int w, int h, int max_threads) {
  return mic_decompress_parallel(c, n, o, w, h, max_threads);
}

}  // extern "C"
