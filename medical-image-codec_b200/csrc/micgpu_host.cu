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
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/micgpu.h"
#include "mic_device.cuh"

using namespace micgpu;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) return fail(MICGPU_E_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      p = nullptr;
      return fail(MICGPU_E_ALLOC, "cudaMalloc(%zu): %s", want, cudaGetErrorString(e));
    }
    cap = want;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

int nstates_index(unsigned n) { return n == 1 ? 0 : n == 2 ? 1 : n == 4 ? 2 : 3; }

struct AnsPlan {
  int nstates = 0, max_log = 0, mode = 0, slots = 0, grid = 0;
};

struct TemporalGroup {
  unsigned long long out_off, fpx;
  int nframes;
};

}  // namespace

struct micgpu_decoder {
  int device = 0;
  int sm_count = 148;
  size_t smem_optin = 227 * 1024;
  std::mutex mu;
  std::vector<MicUnit> units;
  std::vector<int> lists[4];
  std::vector<int> spatial;
  std::vector<TemporalGroup> temporal;
  AnsPlan ans[4];
  unsigned long long sym_total = 0, tab_total = 0, d_total = 0, m_total = 0, out_need = 0;
  int max_log_all = 5, max_w = 1, max_h = 1;
  int k1_grid = 1;
  unsigned long long k1_stride = 0;
  bool committed = false;
  int launches = 0;
  DevBuf d_units, d_list, d_tabA, d_tabS, d_states, d_D, d_M, d_k1, d_comp, d_out;
  MicUnit* h_units = nullptr;   // pinned staging copy
  size_t h_units_cap = 0;
  cudaStream_t stream = nullptr;  // used by the host-buffer convenience calls
  // optional per-kernel timing (CUDA events on the launch stream)
  bool profiling = false;
  std::vector<cudaEvent_t> ev;
  std::vector<std::string> ev_names;
  int ev_used = 0;

  ~micgpu_decoder() {
    cudaSetDevice(device);
    d_units.release(); d_list.release(); d_tabA.release(); d_tabS.release(); d_states.release();
    d_D.release(); d_M.release(); d_k1.release(); d_comp.release(); d_out.release();
    if (h_units) cudaFreeHost(h_units);
    if (stream) cudaStreamDestroy(stream);
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
  for (auto& l : d->lists) l.clear();
  d->spatial.clear();
  for (size_t i = 0; i < d->units.size(); i++) {
    MicUnit& u = d->units[i];
    const unsigned long long px = (unsigned long long)u.width * u.height;
    // symbol capacity: exact for N-state streams; for 1-state the stream is bounded by the
    // RLE worst case (every pixel escaped, plus run headers)
    unsigned long long cap = u.nstates > 1 ? u.count : 2 * px + px / 64 + 4096;
    if (cap > 0xFFFFFFF0ull) { u.status = MIC_E_UNSUPPORTED; cap = 0; }
    if (u.nstates > 1 && cap > 3 * px + 4096) { u.status = MIC_E_SIZE; cap = 0; }   // count cannot exceed the RLE worst case
    u.sym_cap = (unsigned)cap;
    u.sym_off = d->sym_total;
    d->sym_total += (cap + 15) & ~15ull;
    u.tab_off = d->tab_total;
    d->tab_total += 1ull << u.table_log;
    d->max_log_all = std::max(d->max_log_all, (int)u.table_log);
    if (u.kind == MIC_KIND_SPATIAL) {
      u.wp = (u.width + 8 + 63) & ~63u;   // rows start on 128 B lines;   // room for the 0..7 pixel row phase (mic_unit.h align0)
      u.d_off = d->d_total;
      d->d_total += (unsigned long long)u.wp * u.height;
      u.m_off = d->m_total;
      d->m_total += (unsigned long long)(u.wp >> 5) * u.height;
      d->spatial.push_back((int)i);
      d->max_w = std::max(d->max_w, (int)u.width);
      d->max_h = std::max(d->max_h, (int)u.height);
    } else {
      u.wp = 0; u.d_off = 0; u.m_off = 0;
    }
    d->lists[nstates_index(u.nstates)].push_back((int)i);
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
    auto fit = [&](int mode) {
      size_t per = ans_decode_smem_bytes(ml, mode, 1);
      return (int)std::min<size_t>(budget / per, 4096);
    };
    int mode = 0;
    if (fit(0) < std::min(need_per_sm, slots_max)) mode = ml <= 15 && fit(1) > fit(0) ? 1 : 0;
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
  int g1 = std::min<int>((int)d->units.size(), d->sm_count * 8);
  const unsigned long long k1_budget = 256ull << 20;
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
  if ((rc = d->d_k1.ensure((size_t)d->k1_grid * d->k1_stride))) return rc;
  if (nu > d->h_units_cap) {
    if (d->h_units) cudaFreeHost(d->h_units);
    d->h_units = nullptr;
    CUDA_TRY(cudaMallocHost(&d->h_units, (nu + nu / 4 + 16) * sizeof(MicUnit)));
    d->h_units_cap = nu + nu / 4 + 16;
  }
  // unit index lists: [N=1 | N=2 | N=4 | N=8] then the spatial list
  std::vector<int> flat;
  flat.reserve(2 * nu);
  for (auto& l : d->lists) flat.insert(flat.end(), l.begin(), l.end());
  flat.insert(flat.end(), d->spatial.begin(), d->spatial.end());
  if (!flat.empty()) CUDA_TRY(cudaMemcpy(d->d_list.p, flat.data(), flat.size() * sizeof(int), cudaMemcpyHostToDevice));
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

int run_device_locked(micgpu_decoder* d, const void* d_comp, size_t comp_bytes, void* d_out, size_t out_elems, cudaStream_t st) {
  if (!d->committed) return fail(MICGPU_E_HEADER, "decoder plan not committed");
  if (out_elems < d->out_need) return fail(MICGPU_E_SIZE, "output buffer holds %zu elements, plan needs %llu", out_elems, d->out_need);
  for (const MicUnit& u : d->units)
    if (u.comp_off + u.comp_len > comp_bytes) return fail(MICGPU_E_HEADER, "unit extends past the compressed buffer");
  CUDA_TRY(cudaSetDevice(d->device));
  d->launches = 0;
  const int nu = (int)d->units.size();
  if (!nu) return 0;
  for (MicUnit& u : d->units)
    u.align0 = (unsigned)(((reinterpret_cast<uintptr_t>(d_out) >> 1) + u.out_off) & 7u);
  memcpy(d->h_units, d->units.data(), nu * sizeof(MicUnit));
  CUDA_TRY(cudaMemcpyAsync(d->d_units.p, d->h_units, nu * sizeof(MicUnit), cudaMemcpyHostToDevice, st));
  if (d->m_total) CUDA_TRY(cudaMemsetAsync(d->d_M.p, 0, d->m_total * sizeof(uint32_t), st));
  MicUnit* du = (MicUnit*)d->d_units.p;
  const uint8_t* comp = (const uint8_t*)d_comp;
  d->ev_used = 0;
  prof_mark(d, "k_build_tables", st);
  launch_build_tables(du, nu, comp, (uint32_t*)d->d_tabA.p, (uint16_t*)d->d_tabS.p, (uint8_t*)d->d_k1.p, d->k1_stride,
                      d->max_log_all, d->k1_grid, st);
  d->launches++;
  const int* dl = (const int*)d->d_list.p;
  size_t loff = 0;
  for (int g = 0; g < 4; g++) {
    const int n = (int)d->lists[g].size();
    if (n) {
      const AnsPlan& a = d->ans[g];
      static const char* NAMES[4] = {"k_ans_decode<1>", "k_ans_decode<2>", "k_ans_decode<4>", "k_ans_decode<8>"};
      prof_mark(d, NAMES[g], st);
      launch_ans_decode(du, dl + loff, n, a.nstates, comp, (const uint32_t*)d->d_tabA.p, (uint16_t*)d->d_states.p, a.max_log,
                        a.mode, a.slots, a.grid, st);
      d->launches++;
    }
    loff += n;
  }
  prof_mark(d, "k_rle_expand", st);
  launch_rle_expand(du, nu, (const uint16_t*)d->d_states.p, (const uint16_t*)d->d_tabS.p, (uint16_t*)d->d_D.p,
                    (uint32_t*)d->d_M.p, (uint16_t*)d_out, d->max_log_all, std::min(nu, d->sm_count * 8), st);
  d->launches++;
  if (!d->spatial.empty()) {
    prof_mark(d, "k_delta_wavefront", st);
    launch_delta_wavefront(du, dl + loff, (int)d->spatial.size(), (const uint16_t*)d->d_D.p, (const uint32_t*)d->d_M.p,
                           (uint16_t*)d_out, d->max_w, d->max_h, st);
    d->launches++;
  }
  for (const TemporalGroup& t : d->temporal) {
    prof_mark(d, "k_temporal_accumulate", st);
    launch_temporal_accumulate((uint16_t*)d_out + t.out_off, t.fpx, t.nframes, d->sm_count, st);
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
                    Mic2Header& mh) {
  int rc = parse_mic2(p, len, mh);
  if (rc) return rc;
  const unsigned long long fpx = (unsigned long long)mh.w * mh.h;
  const int nf = last_frame < 0 ? mh.n : std::min(mh.n, last_frame + 1);
  for (int i = 0; i < nf; i++) {
    const size_t o = rd32(p + 20 + (size_t)i * 8), l = rd32(p + 24 + (size_t)i * 8);
    if (mh.data_off + o + l > len) return fail(MICGPU_E_HEADER, "MIC2: frame %d data extends beyond file", i);
    const bool residual = mh.temporal && i > 0;
    // residual frames expand to exactly w*h ZigZag words (multiframecompress.go:165-175)
    add_unit_locked(d, p + mh.data_off + o, l, comp_off + mh.data_off + o, residual ? MIC_KIND_RLE : MIC_KIND_SPATIAL,
                    residual ? (uint32_t)fpx : (uint32_t)mh.w, residual ? 1u : (uint32_t)mh.h, out_off + (uint64_t)i * fpx);
  }
  if (mh.temporal && nf > 1) d->temporal.push_back(TemporalGroup{out_off, fpx, nf});
  return 0;
}

// ---- default per-device contexts for the one-shot calls ---------------------
std::mutex g_mu;
micgpu_decoder* g_default[64] = {nullptr};

micgpu_decoder* default_decoder(int dev) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (dev < 0 || dev >= 64) return nullptr;
  if (!g_default[dev]) g_default[dev] = micgpu_decoder_create(dev);
  return g_default[dev];
}

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  return dev;
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

const char* micgpu_last_error(void) { return g_err.c_str(); }

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

void micgpu_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& d : g_default) {
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

int micgpu_decoder_add_pics(micgpu_decoder* d, const uint8_t* pics, size_t len, uint64_t comp_off, uint64_t out_off,
                            int* width, int* height) {
  if (!d || !pics) return fail(MICGPU_E_HEADER, "null argument");
  std::lock_guard<std::mutex> lk(d->mu);
  return add_pics_locked(d, pics, len, comp_off, out_off, width, height);
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
int micgpu_pics_decompress_batch(int n, const uint8_t* const* blobs, const size_t* lens, uint16_t* const* outs,
                                 const size_t* caps, int* status) {
  if (n <= 0) return 0;
  micgpu_decoder* d = default_decoder(current_device());
  if (!d) return MICGPU_E_CUDA;
  std::lock_guard<std::mutex> lk(d->mu);
  d->units.clear();
  d->temporal.clear();
  d->out_need = 0;
  // layout: blobs back to back (64-byte aligned) in one device buffer; outputs back to back
  std::vector<uint64_t> coff(n), ooff(n);
  std::vector<int> first_unit(n + 1, 0), hdr_rc(n, 0);
  uint64_t ctot = 0, otot = 0;
  for (int i = 0; i < n; i++) {
    coff[i] = ctot;
    ooff[i] = otot;
    first_unit[i] = (int)d->units.size();
    int w = 0, h = 0;
    hdr_rc[i] = add_pics_locked(d, blobs[i], lens[i], ctot, otot, &w, &h);
    if (!hdr_rc[i] && (size_t)w * h > caps[i]) {
      hdr_rc[i] = fail(MICGPU_E_SIZE, "image %d: output buffer too small", i);
      d->units.resize(first_unit[i]);
    }
    if (!hdr_rc[i]) otot += (uint64_t)w * h;
    else d->units.resize(first_unit[i]);
    ctot += (lens[i] + 63) & ~(size_t)63;
  }
  first_unit[n] = (int)d->units.size();
  int rc = plan_commit(d);
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(d->device));
  if ((rc = d->d_comp.ensure(ctot + 256))) return rc;
  if ((rc = d->d_out.ensure(std::max<uint64_t>(otot, 1) * sizeof(uint16_t)))) return rc;
  for (int i = 0; i < n; i++)
    if (!hdr_rc[i]) CUDA_TRY(cudaMemcpyAsync((uint8_t*)d->d_comp.p + coff[i], blobs[i], lens[i], cudaMemcpyHostToDevice, d->stream));
  if ((rc = run_device_locked(d, d->d_comp.p, ctot, d->d_out.p, otot, d->stream))) return rc;
  for (int i = 0; i < n; i++) {
    if (hdr_rc[i]) continue;
    const uint64_t px = (i + 1 < n ? ooff[i + 1] : otot) - ooff[i];
    CUDA_TRY(cudaMemcpyAsync(outs[i], (uint16_t*)d->d_out.p + ooff[i], px * sizeof(uint16_t), cudaMemcpyDeviceToHost, d->stream));
  }
  std::vector<int> ust(d->units.size());
  unit_status_locked(d, ust.data(), (int)ust.size(), d->stream);
  int first = 0;
  for (int i = 0; i < n; i++) {
    int s = hdr_rc[i];
    for (int u = first_unit[i]; !s && u < first_unit[i + 1]; u++) s = ust[u];   // first failing strip wins (parallelstrips.go:324-328)
    if (status) status[i] = s;
    if (!first && s) first = s;
  }
  return first;
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
  d->out_need = 0;
  add_unit_locked(d, frame, len, 0, MIC_KIND_SPATIAL, (uint32_t)width, (uint32_t)height, 0);
  int rc = plan_commit(d);
  if (rc) return rc;
  return run_host_locked(d, frame, len, pixels_out, (size_t)width * height);
}

static int mic2_decode(const uint8_t* mic2, size_t len, int last_frame, bool only_last, uint16_t* out, size_t cap_px, int* width,
                       int* height, int* frames, int* temporal) {
  if (!mic2 || !out) return fail(MICGPU_E_HEADER, "null argument");
  micgpu_decoder* d = default_decoder(current_device());
  if (!d) return MICGPU_E_CUDA;
  std::lock_guard<std::mutex> lk(d->mu);
  d->units.clear();
  d->temporal.clear();
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
    if (mh.data_off + o + l > len) return fail(MICGPU_E_HEADER, "MIC2: frame %d data extends beyond file", last_frame);
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

int micgpu_mic2_decompress(const uint8_t* mic2, size_t len, uint16_t* frames_out, size_t cap_px, int* width, int* height,
                           int* frames, int* temporal) {
  return mic2_decode(mic2, len, -1, false, frames_out, cap_px, width, height, frames, temporal);
}

int micgpu_mic2_decompress_frame(const uint8_t* mic2, size_t len, int frame_idx, uint16_t* pixels_out, size_t cap_px, int* width,
                                 int* height) {
  return mic2_decode(mic2, len, frame_idx, true, pixels_out, cap_px, width, height, nullptr, nullptr);
}

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
