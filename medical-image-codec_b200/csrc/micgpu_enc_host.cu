// micgpu_enc_host.cu -- host side of the encode direction: unit planning, the 8->4->2->1 FSE ladder
// (multiframecompress.go:15-93), container assembly (parallelstrips.go:55-123, multiframe.go:49-92) and the C ABI.
#include <algorithm>
#include <mutex>
#include <vector>

#include "host_common.h"
#include "mic_device.cuh"
#include "mic_enc.h"

using namespace micgpu;
using namespace micgpu_host;

struct micgpu_encoder {
  int device = 0;
  int sm_count = 148;
  std::mutex mu;
  std::vector<MicEncUnit> units;
  unsigned long long src_total = 0;
  DevBuf d_units, d_list, d_src, d_V, d_S, d_segs, d_T, d_tab, d_tt, d_hdr, d_frames, d_k6;
  MicEncUnit* h_units = nullptr;
  size_t h_units_cap = 0;
  cudaStream_t stream = nullptr;
  int launches = 0;
  ~micgpu_encoder() {
    cudaSetDevice(device);
    for (DevBuf* b : {&d_units, &d_list, &d_src, &d_V, &d_S, &d_segs, &d_T, &d_tab, &d_tt, &d_hdr, &d_frames, &d_k6}) b->release();
    if (h_units) cudaFreeHost(h_units);
    if (stream) cudaStreamDestroy(stream);
  }
};

namespace {

std::mutex g_enc_mu;
micgpu_encoder* g_enc[64] = {nullptr};

micgpu_encoder* default_encoder(int dev) {
  std::lock_guard<std::mutex> lk(g_enc_mu);
  if (dev < 0 || dev >= 64) return nullptr;
  if (!g_enc[dev]) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || dev >= n) {
      fail(MICGPU_E_CUDA, "no CUDA device available (libmicgpu has no CPU fallback)");
      return nullptr;
    }
    cudaSetDevice(dev);
    micgpu_encoder* e = new micgpu_encoder();
    e->device = dev;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) == cudaSuccess) e->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete e;
      fail(MICGPU_E_CUDA, "cudaStreamCreate failed");
      return nullptr;
    }
    g_enc[dev] = e;
  }
  return g_enc[dev];
}

int bit_len(unsigned v) { int n = 0; while (v) { n++; v >>= 1; } return n; }

// Append one unit; sizes every scratch region from worst-case bounds.
int enc_add_unit(micgpu_encoder* e, int kind, unsigned long long src_off, unsigned width, unsigned height, unsigned max_value, int nstates) {
  MicEncUnit u;
  memset(&u, 0, sizeof u);
  u.kind = (unsigned)kind;
  u.src_off = src_off;
  u.width = width;
  u.height = height;
  u.max_value = max_value;
  u.nstates = (unsigned)nstates;
  e->units.push_back(u);
  return (int)e->units.size() - 1;
}

struct EncTotals { unsigned long long v = 0, s = 0, seg = 0, t = 0, tab = 0, tt = 0, hdr = 0, out = 0; };

void enc_plan(micgpu_encoder* e, EncTotals& T) {
  for (MicEncUnit& u : e->units) {
    const unsigned long long n_in = (unsigned long long)u.width * u.height;
    // V: one symbol per pixel, two when escaped, plus maxValue (spatial); the given stream otherwise
    const unsigned long long vcap = u.kind == MIC_ENC_SPATIAL ? 2 * n_in + 1 : n_in;
    const unsigned long long scap = vcap + vcap / 4 + 16;     // every diff piece adds one header word, same-runs never expand
    const int depth = u.kind == MIC_ENC_SPATIAL ? bit_len(u.max_value) : 16;   // RLE kind: the two length words can be any u16
    const unsigned long long symcap = 1ull << std::max(depth, 1);
    const int lmax = std::min(16, std::max(13, depth + 1));
    u.v_cap = (unsigned)std::min<unsigned long long>(vcap, 0xFFFFFFF0ull);
    u.s_cap = (unsigned)std::min<unsigned long long>(scap, 0xFFFFFFF0ull);
    u.v_off = T.v; T.v += (vcap + 15) & ~15ull;
    u.s_off = T.s; T.s += (scap + 15) & ~15ull;
    u.seg_off = T.seg; T.seg += vcap / 2 + 8;
    u.t_off = T.t; T.t += scap + 16;
    u.tab_off = T.tab; T.tab += 1ull << lmax;
    u.tt_off = T.tt; T.tt += symcap;
    const unsigned long long hcap = symcap * 2 + 64;
    u.hdr_off = T.hdr; T.hdr += (hcap + 15) & ~15ull;
    const unsigned long long ocap = 6 + hcap + 2 * scap + 128;
    u.out_cap = (unsigned)std::min<unsigned long long>(ocap, 0xFFFFFFF0ull);
    u.out_off = T.out; T.out += (ocap + 63) & ~63ull;
    u.status = (vcap > 0x7FFFFFF0ull) ? MIC_ENC_UNSUPPORTED : MIC_ENC_OK;
    if (u.kind == MIC_ENC_SPATIAL && (u.width == 0 || u.height == 0)) u.status = MIC_ENC_UNSUPPORTED;
  }
}

// Runs the whole encode pipeline for e->units whose sources sit in d_src (u16 elements). On return h_units holds the
// device-filled fields and frames are in e->d_frames at out_off.
int enc_run(micgpu_encoder* e, const uint16_t* d_src) {
  CUDA_TRY(cudaSetDevice(e->device));
  const int nu = (int)e->units.size();
  e->launches = 0;
  if (!nu) return 0;
  EncTotals T;
  enc_plan(e, T);
  int rc;
  if ((rc = e->d_units.ensure((size_t)nu * sizeof(MicEncUnit)))) return rc;
  if ((rc = e->d_list.ensure((size_t)nu * sizeof(int) + 64))) return rc;
  if ((rc = e->d_V.ensure((T.v + 64) * 2))) return rc;
  if ((rc = e->d_S.ensure((T.s + 64) * 2))) return rc;
  if ((rc = e->d_segs.ensure((T.seg + 64) * 8))) return rc;
  if ((rc = e->d_T.ensure((T.t + 64) * 4))) return rc;
  if ((rc = e->d_tab.ensure((T.tab + 64) * 2))) return rc;
  if ((rc = e->d_tt.ensure((T.tt + 64) * 8))) return rc;
  if ((rc = e->d_hdr.ensure(T.hdr + 64))) return rc;
  if ((rc = e->d_frames.ensure(T.out + 64))) return rc;
  const int grid = std::min(nu, e->sm_count * 4);
  if ((rc = e->d_k6.ensure((size_t)grid * enc_tables_scratch_per_cta()))) return rc;
  if ((size_t)nu > e->h_units_cap) {
    if (e->h_units) cudaFreeHost(e->h_units);
    e->h_units = nullptr;
    CUDA_TRY(cudaMallocHost(&e->h_units, ((size_t)nu + nu / 4 + 16) * sizeof(MicEncUnit)));
    e->h_units_cap = (size_t)nu + nu / 4 + 16;
  }
  cudaStream_t st = e->stream;
  memcpy(e->h_units, e->units.data(), (size_t)nu * sizeof(MicEncUnit));
  CUDA_TRY(cudaMemcpyAsync(e->d_units.p, e->h_units, (size_t)nu * sizeof(MicEncUnit), cudaMemcpyHostToDevice, st));
  MicEncUnit* du = (MicEncUnit*)e->d_units.p;
  launch_enc_delta_rle(du, nu, d_src, (uint16_t*)e->d_V.p, (uint32_t*)e->d_segs.p, (uint16_t*)e->d_S.p, grid, st);
  launch_enc_tables(du, nu, (const uint16_t*)e->d_S.p, (uint8_t*)e->d_k6.p, (uint16_t*)e->d_tab.p, (uint2*)e->d_tt.p, (uint8_t*)e->d_hdr.p, grid, st);
  e->launches += 5;
  // FSE tiers: a unit rejected at its tier (ErrIncompressible / ErrUseRLE / any error) is retried one tier down, like
  // CompressSingleFrame8State -> 4State -> 2State -> FSECompressU16 (multiframecompress.go:67-93)
  std::vector<int> pending(nu);
  for (int i = 0; i < nu; i++) pending[i] = i;
  for (int round = 0; round < 4 && !pending.empty(); round++) {
    static const int NS[4] = {8, 4, 2, 1};
    std::vector<int> lists[4];
    for (int i : pending) {
      const unsigned n = e->units[i].nstates;
      lists[n == 8 ? 0 : n == 4 ? 1 : n == 2 ? 2 : 3].push_back(i);
    }
    std::vector<int> flat;
    for (auto& l : lists) flat.insert(flat.end(), l.begin(), l.end());
    CUDA_TRY(cudaMemcpyAsync(e->d_list.p, flat.data(), flat.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    size_t off = 0;
    for (int g = 0; g < 4; g++) {
      if (!lists[g].empty()) {
        launch_enc_ans(du, (const int*)e->d_list.p + off, (int)lists[g].size(), NS[g], (const uint16_t*)e->d_S.p, (const uint16_t*)e->d_tab.p,
                       (const uint2*)e->d_tt.p, (uint32_t*)e->d_T.p, e->sm_count, st);
        e->launches++;
      }
      off += lists[g].size();
    }
    launch_enc_pack(du, nu, (const uint32_t*)e->d_T.p, (const uint8_t*)e->d_hdr.p, (uint8_t*)e->d_frames.p, grid, st);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(e->h_units, e->d_units.p, (size_t)nu * sizeof(MicEncUnit), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    // which units fall to the next tier?
    std::vector<int> next;
    for (int i : pending) {
      const MicEncUnit& r = e->h_units[i];
      if (r.status != MIC_ENC_OK && r.status != MIC_ENC_CAPACITY && r.status != MIC_ENC_UNSUPPORTED && e->units[i].nstates > 1 &&
          r.table_log != 0) {   // table_log == 0: the shared front half (histogram / normalisation) failed -> every tier fails the same way
        e->units[i] = r;
        e->units[i].nstates = e->units[i].nstates / 2;
        e->units[i].status = MIC_ENC_OK;
        next.push_back(i);
      }
    }
    if (next.empty()) break;
    // re-upload only the retried descriptors; finished units must not be packed again
    for (int i = 0; i < nu; i++)
      if (std::find(next.begin(), next.end(), i) == next.end()) { e->units[i] = e->h_units[i]; if (e->units[i].status == MIC_ENC_OK) e->units[i].status = 1; }
    memcpy(e->h_units, e->units.data(), (size_t)nu * sizeof(MicEncUnit));
    CUDA_TRY(cudaMemcpyAsync(e->d_units.p, e->h_units, (size_t)nu * sizeof(MicEncUnit), cudaMemcpyHostToDevice, st));
    pending.swap(next);
    if (round == 3) break;
  }
  // final statuses: parked (status 1) units were successes
  CUDA_TRY(cudaMemcpyAsync(e->h_units, e->d_units.p, (size_t)nu * sizeof(MicEncUnit), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  for (int i = 0; i < nu; i++)
    if (e->h_units[i].status == 1) e->h_units[i].status = MIC_ENC_OK;
  return 0;
}

int enc_status_to_rc(int st) {
  switch (st) {
    case MIC_ENC_OK: return 0;
    case MIC_ENC_INCOMPRESSIBLE: return fail(MICGPU_E_INCOMPRESSIBLE, "input is not compressible");
    case MIC_ENC_USE_RLE: return fail(MICGPU_E_USE_RLE, "input is single value repeated");
    case MIC_ENC_CAPACITY: return fail(MICGPU_E_SIZE, "encode scratch capacity exceeded");
    case MIC_ENC_UNSUPPORTED: return fail(MICGPU_E_UNSUPPORTED, "maxValue of fewer than 4 bits (or zero) is not codable");
    default: return fail(MICGPU_E_INTERNAL, "FSE table construction failed");
  }
}

}  // namespace

extern "C" {

// DeltaRleCompressU16.Compress (deltarlecompressu16.go:24-68): the RLE symbol stream of one image (stage API, used by tests)
int micgpu_delta_rle_compress(const uint16_t* pixels, int width, int height, uint16_t max_value, uint16_t* out, size_t cap, size_t* out_len) {
  if (!pixels || !out || width <= 0 || height <= 0) return fail(MICGPU_E_HEADER, "bad argument");
  micgpu_encoder* e = default_encoder(current_device());
  if (!e) return MICGPU_E_CUDA;
  std::lock_guard<std::mutex> lk(e->mu);
  e->units.clear();
  enc_add_unit(e, MIC_ENC_SPATIAL, 0, (unsigned)width, (unsigned)height, max_value, 2);
  const size_t npx = (size_t)width * height;
  int rc;
  CUDA_TRY(cudaSetDevice(e->device));
  if ((rc = e->d_src.ensure(npx * 2 + 64))) return rc;
  CUDA_TRY(cudaMemcpyAsync(e->d_src.p, pixels, npx * 2, cudaMemcpyHostToDevice, e->stream));
  if ((rc = enc_run(e, (const uint16_t*)e->d_src.p))) return rc;
  const MicEncUnit& r = e->h_units[0];
  if (r.s_len == 0 && r.status != MIC_ENC_OK) return enc_status_to_rc(r.status);
  if (r.s_len > cap) return fail(MICGPU_E_SIZE, "output buffer too small");
  CUDA_TRY(cudaMemcpy(out, (uint16_t*)e->d_S.p + r.s_off, (size_t)r.s_len * 2, cudaMemcpyDeviceToHost));
  if (out_len) *out_len = r.s_len;
  return 0;
}

// RleCompressU16.Init(len,1,maxValue) + Compress (rlecompressu16.go:15-93) (stage API)
int micgpu_rle_compress(const uint16_t* in, size_t n, uint16_t max_value, uint16_t* out, size_t cap, size_t* out_len) {
  if (!in || !out || n == 0 || n > 0x7FFFFFF0u) return fail(MICGPU_E_HEADER, "bad argument");
  micgpu_encoder* e = default_encoder(current_device());
  if (!e) return MICGPU_E_CUDA;
  std::lock_guard<std::mutex> lk(e->mu);
  e->units.clear();
  enc_add_unit(e, MIC_ENC_RLE, 0, (unsigned)n, 1, max_value, 2);
  int rc;
  CUDA_TRY(cudaSetDevice(e->device));
  if ((rc = e->d_src.ensure(n * 2 + 64))) return rc;
  CUDA_TRY(cudaMemcpyAsync(e->d_src.p, in, n * 2, cudaMemcpyHostToDevice, e->stream));
  if ((rc = enc_run(e, (const uint16_t*)e->d_src.p))) return rc;
  const MicEncUnit& r = e->h_units[0];
  if (r.s_len == 0 && r.status != MIC_ENC_OK) return enc_status_to_rc(r.status);
  if (r.s_len > cap) return fail(MICGPU_E_SIZE, "output buffer too small");
  CUDA_TRY(cudaMemcpy(out, (uint16_t*)e->d_S.p + r.s_off, (size_t)r.s_len * 2, cudaMemcpyDeviceToHost));
  if (out_len) *out_len = r.s_len;
  return 0;
}

// CompressSingleFrame / 4State / 8State (multiframecompress.go:15-93); nstates 1 = plain FSECompressU16 (fseu16_test.go:679)
int micgpu_compress_single_frame(const uint16_t* pixels, int width, int height, uint16_t max_value, int nstates, uint8_t* out, size_t cap,
                                 size_t* out_len) {
  if (!pixels || !out || width <= 0 || height <= 0) return fail(MICGPU_E_HEADER, "bad argument");
  if (nstates != 1 && nstates != 2 && nstates != 4 && nstates != 8) return fail(MICGPU_E_HEADER, "nstates must be 1, 2, 4 or 8");
  micgpu_encoder* e = default_encoder(current_device());
  if (!e) return MICGPU_E_CUDA;
  std::lock_guard<std::mutex> lk(e->mu);
  e->units.clear();
  enc_add_unit(e, MIC_ENC_SPATIAL, 0, (unsigned)width, (unsigned)height, max_value, nstates);
  const size_t npx = (size_t)width * height;
  int rc;
  CUDA_TRY(cudaSetDevice(e->device));
  if ((rc = e->d_src.ensure(npx * 2 + 64))) return rc;
  CUDA_TRY(cudaMemcpyAsync(e->d_src.p, pixels, npx * 2, cudaMemcpyHostToDevice, e->stream));
  if ((rc = enc_run(e, (const uint16_t*)e->d_src.p))) return rc;
  const MicEncUnit& r = e->h_units[0];
  if (r.status != MIC_ENC_OK) return enc_status_to_rc(r.status);
  if (r.frame_len > cap) return fail(MICGPU_E_SIZE, "output buffer too small");
  CUDA_TRY(cudaMemcpy(out, (uint8_t*)e->d_frames.p + r.out_off, r.frame_len, cudaMemcpyDeviceToHost));
  if (out_len) *out_len = r.frame_len;
  return 0;
}

// CompressParallelStrips / 4State / 8State for n images of the same geometry (parallelstrips.go:55-265)
int micgpu_pics_compress_batch(int n, const uint16_t* const* pixels, int width, int height, const uint16_t* max_values, int num_strips,
                               int nstates, uint8_t* const* outs, const size_t* caps, size_t* out_lens, int* status) {
  if (n <= 0) return 0;
  if (!pixels || !outs || width <= 0 || height <= 0 || num_strips <= 0) return fail(MICGPU_E_HEADER, "bad argument");
  if (nstates != 2 && nstates != 4 && nstates != 8) return fail(MICGPU_E_HEADER, "nstates must be 2, 4 or 8");
  micgpu_encoder* e = default_encoder(current_device());
  if (!e) return MICGPU_E_CUDA;
  std::lock_guard<std::mutex> lk(e->mu);
  // strip geometry (parallelstrips.go:62-72)
  int ns = std::min(num_strips, height);
  if (ns < 1) ns = 1;
  const int strip_h = (height + ns - 1) / ns;
  const int actual = (height + strip_h - 1) / strip_h;
  const size_t npx = (size_t)width * height;
  e->units.clear();
  for (int i = 0; i < n; i++)
    for (int s = 0; s < actual; s++) {
      const int y0 = s * strip_h, y1 = std::min(height, y0 + strip_h);
      enc_add_unit(e, MIC_ENC_SPATIAL, (unsigned long long)i * npx + (unsigned long long)y0 * width, (unsigned)width, (unsigned)(y1 - y0),
                   max_values[i], nstates);
    }
  int rc;
  CUDA_TRY(cudaSetDevice(e->device));
  if ((rc = e->d_src.ensure((size_t)n * npx * 2 + 64))) return rc;
  for (int i = 0; i < n; i++)
    CUDA_TRY(cudaMemcpyAsync((uint16_t*)e->d_src.p + (size_t)i * npx, pixels[i], npx * 2, cudaMemcpyHostToDevice, e->stream));
  if ((rc = enc_run(e, (const uint16_t*)e->d_src.p))) return rc;
  int first = 0;
  for (int i = 0; i < n; i++) {
    int st = 0;
    size_t total = 20 + (size_t)actual * 8;
    for (int s = 0; s < actual && !st; s++) {
      const MicEncUnit& r = e->h_units[(size_t)i * actual + s];
      if (r.status != MIC_ENC_OK) st = enc_status_to_rc(r.status);   // "parallelstrips: strip %d: ..." first error wins
      total += r.frame_len;
    }
    if (!st && total > caps[i]) st = fail(MICGPU_E_SIZE, "image %d: output buffer too small", i);
    if (!st) {
      uint8_t* o = outs[i];
      auto put32 = [&](size_t off, uint32_t v) { o[off] = (uint8_t)v; o[off + 1] = (uint8_t)(v >> 8); o[off + 2] = (uint8_t)(v >> 16); o[off + 3] = (uint8_t)(v >> 24); };
      memcpy(o, "PICS", 4);
      put32(4, (uint32_t)width); put32(8, (uint32_t)height); put32(12, (uint32_t)actual); put32(16, (uint32_t)strip_h);
      size_t off = 0;
      const size_t hdr = 20 + (size_t)actual * 8;
      for (int s = 0; s < actual; s++) {
        const MicEncUnit& r = e->h_units[(size_t)i * actual + s];
        put32(20 + (size_t)s * 8, (uint32_t)off); put32(24 + (size_t)s * 8, r.frame_len);
        CUDA_TRY(cudaMemcpyAsync(o + hdr + off, (uint8_t*)e->d_frames.p + r.out_off, r.frame_len, cudaMemcpyDeviceToHost, e->stream));
        off += r.frame_len;
      }
      if (out_lens) out_lens[i] = total;
    }
    if (status) status[i] = st;
    if (!first && st) first = st;
  }
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  return first;
}

int micgpu_pics_compress(const uint16_t* pixels, int width, int height, uint16_t max_value, int num_strips, int nstates, uint8_t* out,
                         size_t cap, size_t* out_len) {
  return micgpu_pics_compress_batch(1, &pixels, width, height, &max_value, num_strips, nstates, &out, &cap, out_len, nullptr);
}

// mic_compress_{two,four,eight}_state (ojph/mic_compress_c.h:26-37): the C twin derives maxValue from the pixels
// (ojph/mic_compress_c.c:774-775) and has no fallback ladder: any rejection is an error.
static int twin_compress(const uint16_t* pixels, int width, int height, uint8_t* out, size_t out_cap, size_t* out_len, int nstates) {
  if (!pixels || !out || width <= 0 || height <= 0) return fail(MICGPU_E_HEADER, "bad argument");
  uint16_t mx = 0;
  for (size_t i = 0, n = (size_t)width * height; i < n; i++) mx = std::max(mx, pixels[i]);
  size_t len = 0;
  int rc = micgpu_compress_single_frame(pixels, width, height, mx, nstates, out, out_cap, &len);
  if (rc) return rc;
  const uint8_t want = nstates == 2 ? 0x02 : nstates == 4 ? 0x04 : 0x84;
  if (len < 2 || out[0] != 0xFF || out[1] != want) return fail(MICGPU_E_INCOMPRESSIBLE, "input rejected by the %d-state coder", nstates);
  if (out_len) *out_len = len;
  return 0;
}
int mic_compress_two_state(const uint16_t* p, int w, int h, uint8_t* o, size_t c, size_t* l) { return twin_compress(p, w, h, o, c, l, 2); }
int mic_compress_four_state(const uint16_t* p, int w, int h, uint8_t* o, size_t c, size_t* l) { return twin_compress(p, w, h, o, c, l, 4); }
int mic_compress_eight_state(const uint16_t* p, int w, int h, uint8_t* o, size_t c, size_t* l) { return twin_compress(p, w, h, o, c, l, 8); }

void micgpu_encoder_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_enc_mu);
  for (auto& e : g_enc) { delete e; e = nullptr; }
}

}  // extern "C"
