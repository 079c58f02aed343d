// micgpu_enc_host.cu -- host side of the encode direction: unit planning, the 8->4->2->1 FSE ladder
// (multiframecompress.go:15-93), container assembly (parallelstrips.go:55-123, multiframe.go:49-92) and the C ABI.
#include <algorithm>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
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
  DevBuf d_img, d_planes, d_jobs, d_stats, d_compact, d_wA, d_wB;   // front ends
  MicEncUnit* h_units = nullptr;
  size_t h_units_cap = 0;
  cudaStream_t stream = nullptr;
  int launches = 0;
  ~micgpu_encoder() {
    cudaSetDevice(device);
    for (DevBuf* b : {&d_units, &d_list, &d_src, &d_V, &d_S, &d_segs, &d_T, &d_tab, &d_tt, &d_hdr, &d_frames, &d_k6, &d_img, &d_planes, &d_jobs,
                      &d_stats, &d_compact, &d_wA, &d_wB})
      b->release();
    if (h_units) cudaFreeHost(h_units);
    if (stream) cudaStreamDestroy(stream);
  }
};

namespace {

std::mutex g_enc_mu;
constexpr int ENC_CTX = 8;   // [0] serves the one-shot calls; the batch call pipelines chunks over several of them
micgpu_encoder* g_enc_ctx[64][ENC_CTX] = {};

micgpu_encoder* encoder_ctx(int dev, int slot) {
  std::lock_guard<std::mutex> lk(g_enc_mu);
  if (dev < 0 || dev >= 64 || slot < 0 || slot >= ENC_CTX) return nullptr;
  micgpu_encoder*& ref = g_enc_ctx[dev][slot];
  if (!ref) {
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
    ref = e;
  }
  return ref;
}

micgpu_encoder* default_encoder(int dev) { return encoder_ctx(dev, 0); }

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
using EncHook = std::function<void(MicEncUnit*, cudaStream_t)>;

int enc_run(micgpu_encoder* e, const uint16_t* d_src, const EncHook& after_upload = EncHook()) {
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
  // MICGPU_DEBUG_SYNC=1: synchronise after every launch group and name the one that faulted
  static const bool dbg = [] { const char* v = getenv("MICGPU_DEBUG_SYNC"); return v && v[0] == '1'; }();
  auto chk = [&](const char* what) -> int {
    if (!dbg) return 0;
    cudaError_t ce = cudaStreamSynchronize(st);
    if (ce == cudaSuccess) ce = cudaGetLastError();
    if (ce != cudaSuccess) return fail(MICGPU_E_CUDA, "%s: %s", what, cudaGetErrorString(ce));
    return 0;
  };
  memcpy(e->h_units, e->units.data(), (size_t)nu * sizeof(MicEncUnit));
  CUDA_TRY(cudaMemcpyAsync(e->d_units.p, e->h_units, (size_t)nu * sizeof(MicEncUnit), cudaMemcpyHostToDevice, st));
  MicEncUnit* du = (MicEncUnit*)e->d_units.p;
  if (after_upload) after_upload(du, st);
  if (dbg) {   // poison every scratch buffer so that a read of something no kernel wrote misbehaves the same way in every process
    cudaMemsetAsync(e->d_V.p, 0xFF, (T.v + 64) * 2, st); cudaMemsetAsync(e->d_S.p, 0xFF, (T.s + 64) * 2, st);
    cudaMemsetAsync(e->d_segs.p, 0xFF, (T.seg + 64) * 8, st); cudaMemsetAsync(e->d_T.p, 0xFF, (T.t + 64) * 4, st);
    cudaMemsetAsync(e->d_tab.p, 0xFF, (T.tab + 64) * 2, st); cudaMemsetAsync(e->d_tt.p, 0xFF, (T.tt + 64) * 8, st);
    cudaMemsetAsync(e->d_hdr.p, 0xFF, T.hdr + 64, st); cudaMemsetAsync(e->d_frames.p, 0xFF, T.out + 64, st);
  }
  if ((rc = chk("upload"))) return rc;
  launch_enc_delta_rle(du, nu, d_src, (uint16_t*)e->d_V.p, (uint32_t*)e->d_segs.p, (uint16_t*)e->d_S.p, grid, st);
  if ((rc = chk("k_enc_delta / k_enc_rle_*"))) return rc;
  launch_enc_tables(du, nu, (const uint16_t*)e->d_S.p, (uint8_t*)e->d_k6.p, (uint16_t*)e->d_tab.p, (uint2*)e->d_tt.p, (uint8_t*)e->d_hdr.p, grid, st);
  e->launches += 5;
  if ((rc = chk("k_enc_tables"))) return rc;
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
        // the largest table a unit of this list can have (the reservation of enc_plan) decides whether the state tables
        // go to shared memory (k_enc_ans_smem)
        int max_log = 0;
        bool any_rans = false;
        for (int i : lists[g]) {
          const MicEncUnit& u = e->units[i];
          const int depth = u.kind == MIC_ENC_SPATIAL ? bit_len(u.max_value) : 16;
          max_log = std::max(max_log, std::min(16, std::max(13, depth + 1)));
          any_rans |= u.rans != 0;
        }
        launch_enc_ans(du, (const int*)e->d_list.p + off, (int)lists[g].size(), NS[g], (const uint16_t*)e->d_S.p, (const uint16_t*)e->d_tab.p,
                       (const uint2*)e->d_tt.p, (uint32_t*)e->d_T.p, e->sm_count, st, max_log, any_rans);
        e->launches++;
      }
      off += lists[g].size();
    }
    if ((rc = chk("k_enc_ans"))) return rc;
    launch_enc_pack(du, nu, (const uint32_t*)e->d_T.p, (const uint8_t*)e->d_hdr.p, (uint8_t*)e->d_frames.p, grid, st);
    e->launches++;
    if ((rc = chk("k_enc_pack"))) return rc;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(e->h_units, e->d_units.p, (size_t)nu * sizeof(MicEncUnit), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    // which units fall to the next tier?
    std::vector<int> next;
    for (int i : pending) {
      const MicEncUnit& r = e->h_units[i];
      if (r.status != MIC_ENC_OK && r.status != MIC_ENC_CAPACITY && r.status != MIC_ENC_UNSUPPORTED && e->units[i].nstates > 1 &&
          !e->units[i].no_ladder && r.table_log != 0) {   // table_log == 0: the shared front half (histogram / normalisation) failed -> every tier fails the same way
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

int err_code_of(int st) {
  switch (st) {
    case MIC_ENC_OK: return 0;
    case MIC_ENC_INCOMPRESSIBLE: return MICGPU_E_INCOMPRESSIBLE;
    case MIC_ENC_USE_RLE: return MICGPU_E_USE_RLE;
    case MIC_ENC_CAPACITY: return MICGPU_E_SIZE;
    case MIC_ENC_UNSUPPORTED: return MICGPU_E_UNSUPPORTED;
    default: return MICGPU_E_INTERNAL;
  }
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
  const bool rans = nstates == MICGPU_CODER_RANS8;
  if (nstates != 1 && nstates != 2 && nstates != 4 && nstates != 8 && !rans) return fail(MICGPU_E_HEADER, "nstates must be 1, 2, 4, 8 or MICGPU_CODER_RANS8");
  micgpu_encoder* e = default_encoder(current_device());
  if (!e) return MICGPU_E_CUDA;
  std::lock_guard<std::mutex> lk(e->mu);
  e->units.clear();
  const int ui = enc_add_unit(e, MIC_ENC_SPATIAL, 0, (unsigned)width, (unsigned)height, max_value, rans ? 8 : nstates);
  if (rans) { e->units[ui].rans = 1; e->units[ui].no_ladder = 1; }   // RANSCompressU16EightState has no fallback (rans8state.go:31-70)
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

// CompressSingleFrameGrad (multiframecompress.go:111-127): gradient-adaptive Delta + RLE, two-state FSE, one-state fallback
int micgpu_compress_single_frame_grad(const uint16_t* pixels, int width, int height, uint16_t max_value, uint8_t* out, size_t cap,
                                      size_t* out_len) {
  if (!pixels || !out || width <= 0 || height <= 0) return fail(MICGPU_E_HEADER, "bad argument");
  micgpu_encoder* e = default_encoder(current_device());
  if (!e) return MICGPU_E_CUDA;
  std::lock_guard<std::mutex> lk(e->mu);
  e->units.clear();
  const int ui = enc_add_unit(e, MIC_ENC_SPATIAL, 0, (unsigned)width, (unsigned)height, max_value, 2);
  e->units[ui].predictor = 1;
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

// adaptiveStripBoundaries (parallelstripsadaptive.go:214-289) from the per-row costs the GPU summed.  The reference
// accumulates in float64; the sums are integers far below 2^53, so the doubles here take exactly the same values.
static void pica_boundaries(const unsigned long long* cost, int height, int num_strips, std::vector<int>& starts) {
  starts.clear();
  if (num_strips >= height) {
    for (int i = 0; i < height; i++) starts.push_back(i);
    return;
  }
  if (num_strips == 1) { starts.push_back(0); return; }
  std::vector<double> cum((size_t)height + 1, 0.0);
  for (int y = 0; y < height; y++) cum[y + 1] = cum[y] + (double)cost[y];
  const double total = cum[height];
  starts.assign(num_strips, 0);
  if (total == 0) {
    for (int i = 1; i < num_strips; i++) starts[i] = (int)((long long)i * height / num_strips);
    return;
  }
  for (int i = 1; i < num_strips; i++) {
    const double target = total * (double)i / (double)num_strips;
    int lo = starts[i - 1] + 1, hi = height;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cum[mid] < target) lo = mid + 1; else hi = mid;
    }
    if (lo >= height) lo = height - 1;
    starts[i] = lo;
  }
}

// CompressParallelStripsAdaptive (parallelstripsadaptive.go:54-139): content-adaptive strip boundaries, every strip coded
// with both predictors (two units per strip, one launch sequence for all of them), the smaller frame kept (the gradient
// one on a tie), "PICA" container.  starts_out (optional, >= min(num_strips, height) ints) receives the strip start rows.
int micgpu_pica_compress(const uint16_t* pixels, int width, int height, uint16_t max_value, int num_strips, uint8_t* out, size_t cap,
                         size_t* out_len) {
  if (!pixels || !out || width <= 0 || height <= 0 || num_strips <= 0) return fail(MICGPU_E_HEADER, "bad argument");
  micgpu_encoder* e = default_encoder(current_device());
  if (!e) return MICGPU_E_CUDA;
  std::lock_guard<std::mutex> lk(e->mu);
  int ns = std::min(num_strips, height);
  if (ns < 1) ns = 1;
  const size_t npx = (size_t)width * height;
  int rc;
  CUDA_TRY(cudaSetDevice(e->device));
  if ((rc = e->d_src.ensure(npx * 2 + 64))) return rc;
  CUDA_TRY(cudaMemcpyAsync(e->d_src.p, pixels, npx * 2, cudaMemcpyHostToDevice, e->stream));
  std::vector<unsigned long long> cost((size_t)height, 0);
  if (ns > 1 && ns < height) {
    if ((rc = e->d_stats.ensure((size_t)height * sizeof(unsigned long long)))) return rc;
    launch_row_costs((const uint16_t*)e->d_src.p, width, height, (unsigned long long*)e->d_stats.p, e->stream);
    CUDA_TRY(cudaMemcpyAsync(cost.data(), e->d_stats.p, (size_t)height * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
  }
  std::vector<int> starts;
  pica_boundaries(cost.data(), height, ns, starts);
  const int actual = (int)starts.size();
  e->units.clear();
  for (int s = 0; s < actual; s++) {
    const int y0 = starts[s], y1 = s + 1 < actual ? starts[s + 1] : height;
    for (int g = 0; g < 2; g++) {
      // a boundary list may repeat a row (targets beyond the last row collapse onto height-1): such a strip has no rows,
      // the reference then fails in FSE on the one-symbol stream; enc_plan marks a zero-height unit unsupported
      const int ui = enc_add_unit(e, MIC_ENC_SPATIAL, (unsigned long long)y0 * width, (unsigned)width, (unsigned)std::max(0, y1 - y0), max_value, 2);
      e->units[ui].predictor = (unsigned)g;
    }
  }
  if ((rc = enc_run(e, (const uint16_t*)e->d_src.p))) return rc;
  size_t total = 16 + (size_t)actual * 16;
  std::vector<int> pick(actual);
  for (int s = 0; s < actual; s++) {
    const MicEncUnit &a = e->h_units[2 * s], &g = e->h_units[2 * s + 1];
    const bool ga = g.status == MIC_ENC_OK && (a.status != MIC_ENC_OK || g.frame_len <= a.frame_len);
    if (!ga && a.status != MIC_ENC_OK) { enc_status_to_rc(a.status); return fail(err_code_of(a.status), "pica: strip %d: %s", s, std::string(err_slot()).c_str()); }
    pick[s] = 2 * s + (ga ? 1 : 0);
    total += e->h_units[pick[s]].frame_len;
  }
  if (total > cap) return fail(MICGPU_E_SIZE, "output buffer too small");
  auto put = [&](size_t off, uint32_t v) { out[off] = (uint8_t)v; out[off + 1] = (uint8_t)(v >> 8); out[off + 2] = (uint8_t)(v >> 16); out[off + 3] = (uint8_t)(v >> 24); };
  memcpy(out, "PICA", 4);
  put(4, (uint32_t)width); put(8, (uint32_t)height); put(12, (uint32_t)actual);
  const size_t hdr = 16 + (size_t)actual * 16;
  size_t off = 0;
  for (int s = 0; s < actual; s++) {
    const MicEncUnit& r = e->h_units[pick[s]];
    const size_t eo = 16 + (size_t)s * 16;
    put(eo, (uint32_t)starts[s]); put(eo + 4, (uint32_t)off); put(eo + 8, r.frame_len); put(eo + 12, (uint32_t)(pick[s] & 1));
    CUDA_TRY(cudaMemcpyAsync(out + hdr + off, (uint8_t*)e->d_frames.p + r.out_off, r.frame_len, cudaMemcpyDeviceToHost, e->stream));
    off += r.frame_len;
  }
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  if (out_len) *out_len = total;
  return 0;
}

// the strip start rows CompressParallelStripsAdaptive would choose (test / tooling hook for adaptiveStripBoundaries)
int micgpu_pica_boundaries(const uint16_t* pixels, int width, int height, int num_strips, int* starts_out, int* n_out) {
  if (!pixels || !starts_out || width <= 0 || height <= 0 || num_strips <= 0) return fail(MICGPU_E_HEADER, "bad argument");
  micgpu_encoder* e = default_encoder(current_device());
  if (!e) return MICGPU_E_CUDA;
  std::lock_guard<std::mutex> lk(e->mu);
  const int ns = std::max(1, std::min(num_strips, height));
  const size_t npx = (size_t)width * height;
  int rc;
  CUDA_TRY(cudaSetDevice(e->device));
  if ((rc = e->d_src.ensure(npx * 2 + 64))) return rc;
  if ((rc = e->d_stats.ensure((size_t)height * sizeof(unsigned long long)))) return rc;
  CUDA_TRY(cudaMemcpyAsync(e->d_src.p, pixels, npx * 2, cudaMemcpyHostToDevice, e->stream));
  launch_row_costs((const uint16_t*)e->d_src.p, width, height, (unsigned long long*)e->d_stats.p, e->stream);
  std::vector<unsigned long long> cost((size_t)height, 0);
  CUDA_TRY(cudaMemcpyAsync(cost.data(), e->d_stats.p, (size_t)height * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  std::vector<int> starts;
  pica_boundaries(cost.data(), height, ns, starts);
  for (size_t i = 0; i < starts.size(); i++) starts_out[i] = starts[i];
  if (n_out) *n_out = (int)starts.size();
  return 0;
}

// CompressParallelStrips / 4State / 8State for images [i0, i1) of one geometry on one encoder context
// (parallelstrips.go:55-265).  *msg receives the error text (fail() is per thread).
static int pics_compress_range(micgpu_encoder* e, int i0, int i1, const uint16_t* const* pixels, int width, int height,
                               const uint16_t* max_values, int num_strips, int nstates, uint8_t* const* outs, const size_t* caps,
                               size_t* out_lens, int* status, std::string* msg) {
  std::lock_guard<std::mutex> lk(e->mu);
  // an error return first waits for whatever this context still has in flight (copies into the caller's buffers)
  auto done = [&](int rc) { if (rc) { cudaStreamSynchronize(e->stream); if (msg) *msg = err_slot(); } return rc; };
  const int n = i1 - i0;
  // strip geometry (parallelstrips.go:62-72)
  int ns = std::min(num_strips, height);
  if (ns < 1) ns = 1;
  const int strip_h = (height + ns - 1) / ns;
  const int actual = (height + strip_h - 1) / strip_h;
  const size_t npx = (size_t)width * height;
  e->units.clear();
  for (int k = 0; k < n; k++)
    for (int s = 0; s < actual; s++) {
      const int y0 = s * strip_h, y1 = std::min(height, y0 + strip_h);
      enc_add_unit(e, MIC_ENC_SPATIAL, (unsigned long long)k * npx + (unsigned long long)y0 * width, (unsigned)width, (unsigned)(y1 - y0),
                   max_values[i0 + k], nstates);
    }
  int rc;
  if (cudaSetDevice(e->device) != cudaSuccess) return done(fail(MICGPU_E_CUDA, "cudaSetDevice failed"));
  if ((rc = e->d_src.ensure((size_t)n * npx * 2 + 64))) return done(rc);
  // images that are neighbours in host memory travel in one copy
  for (int k = 0; k < n;) {
    int j = k;
    while (j + 1 < n && pixels[i0 + j + 1] == pixels[i0 + j] + npx) j++;
    if (cudaMemcpyAsync((uint16_t*)e->d_src.p + (size_t)k * npx, pixels[i0 + k], (size_t)(j - k + 1) * npx * 2, cudaMemcpyHostToDevice, e->stream) != cudaSuccess)
      return done(fail(MICGPU_E_CUDA, "H2D copy of image %d failed", i0 + k));
    k = j + 1;
  }
  if ((rc = enc_run(e, (const uint16_t*)e->d_src.p))) return done(rc);
  int first = 0;
  for (int k = 0; k < n; k++) {
    const int i = i0 + k;
    int st = 0;
    size_t total = 20 + (size_t)actual * 8;
    for (int s = 0; s < actual && !st; s++) {
      const MicEncUnit& r = e->h_units[(size_t)k * actual + s];
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
        const MicEncUnit& r = e->h_units[(size_t)k * actual + s];
        put32(20 + (size_t)s * 8, (uint32_t)off); put32(24 + (size_t)s * 8, r.frame_len);
        if (cudaMemcpyAsync(o + hdr + off, (uint8_t*)e->d_frames.p + r.out_off, r.frame_len, cudaMemcpyDeviceToHost, e->stream) != cudaSuccess)
          return done(fail(MICGPU_E_CUDA, "D2H copy of image %d failed", i));
        off += r.frame_len;
      }
      if (out_lens) out_lens[i] = total;
    } else if (!first && msg) {
      *msg = err_slot();
    }
    if (status) status[i] = st;
    if (!first && st) first = st;
  }
  if (cudaStreamSynchronize(e->stream) != cudaSuccess) return done(fail(MICGPU_E_CUDA, "encode stream failed"));
  return first;
}

// Batch entry.  Large batches are cut into chunks that worker threads push through several encoder contexts (own stream,
// own scratch): the H2D copy of one chunk, the kernels of another and the D2H copies of a third overlap, and the host
// round trips of the FSE tier ladder (enc_run synchronises to read each tier's verdicts) no longer idle the GPU.
// Measured on 256 radiographs in pinned memory (tools/enc_e2e.py): one context 3 GB/s (pageable) -> 4 contexts x 1 chunk
// 32 GB/s; more, smaller chunks are slower because k_enc_ans costs ~15 ms per chunk whatever its size (serial chains).
int micgpu_pics_compress_batch(int n, const uint16_t* const* pixels, int width, int height, const uint16_t* max_values, int num_strips,
                               int nstates, uint8_t* const* outs, const size_t* caps, size_t* out_lens, int* status) {
  if (n <= 0) return 0;
  if (!pixels || !outs || width <= 0 || height <= 0 || num_strips <= 0) return fail(MICGPU_E_HEADER, "bad argument");
  if (nstates != 2 && nstates != 4 && nstates != 8) return fail(MICGPU_E_HEADER, "nstates must be 2, 4 or 8");
  const int dev = current_device();
  const size_t img_bytes = (size_t)width * height * 2;
  static const int cfg_workers = [] { const char* v = getenv("MICGPU_ENC_CTX"); int k = v ? atoi(v) : 4; return k < 1 ? 1 : (k > ENC_CTX ? ENC_CTX : k); }();
  static const int cfg_per = [] { const char* v = getenv("MICGPU_ENC_CHUNKS_PER_CTX"); int k = v ? atoi(v) : 1; return k < 1 ? 1 : k; }();
  const int workers = (n >= 16 && (size_t)n * img_bytes >= ((size_t)64 << 20)) ? cfg_workers : 1;
  if (workers == 1) {
    micgpu_encoder* e = default_encoder(dev);
    if (!e) return MICGPU_E_CUDA;
    std::string msg;
    const int rc = pics_compress_range(e, 0, n, pixels, width, height, max_values, num_strips, nstates, outs, caps, out_lens, status, &msg);
    if (rc && !msg.empty()) err_slot() = msg;
    return rc;
  }
  micgpu_encoder* ctx[ENC_CTX];
  for (int t = 0; t < workers; t++)
    if (!(ctx[t] = encoder_ctx(dev, t))) return MICGPU_E_CUDA;
  const int chunk = std::max(4, (n + cfg_per * workers - 1) / (cfg_per * workers));
  const int nchunks = (n + chunk - 1) / chunk;
  std::vector<int> rcs(nchunks, 0);
  std::vector<std::string> msgs(nchunks);
  std::vector<std::thread> pool;
  for (int t = 0; t < workers; t++)
    pool.emplace_back([&, t]() {
      cudaSetDevice(dev);
      for (int c = t; c < nchunks; c += workers) {
        const int i0 = c * chunk, i1 = std::min(n, i0 + chunk);
        rcs[c] = pics_compress_range(ctx[t], i0, i1, pixels, width, height, max_values, num_strips, nstates, outs, caps, out_lens, status, &msgs[c]);
      }
    });
  for (auto& th : pool) th.join();
  for (int c = 0; c < nchunks; c++)
    if (rcs[c]) {
      if (!msgs[c].empty()) err_slot() = msgs[c];
      return rcs[c];
    }
  return 0;
}

int micgpu_pics_compress(const uint16_t* pixels, int width, int height, uint16_t max_value, int num_strips, int nstates, uint8_t* out,
                         size_t cap, size_t* out_len) {
  return micgpu_pics_compress_batch(1, &pixels, width, height, &max_value, num_strips, nstates, &out, &cap, out_len, nullptr);
}

// ---- helpers for container assembly -------------------------------------------------------------------
}  // extern "C"

namespace {

void put32(std::vector<uint8_t>& v, uint32_t x) { for (int i = 0; i < 4; i++) v.push_back((uint8_t)(x >> (8 * i))); }
void put64(std::vector<uint8_t>& v, uint64_t x) { for (int i = 0; i < 8; i++) v.push_back((uint8_t)(x >> (8 * i))); }

// Copy the finished frames of units [u0, u1) to the host, compacted on the device first (one D2H instead of one per frame).
int fetch_frames(micgpu_encoder* e, int u0, int u1, std::vector<std::vector<uint8_t>>& out) {
  std::vector<GatherJob> jobs;
  unsigned long long tot = 0;
  for (int i = u0; i < u1; i++) {
    const MicEncUnit& r = e->h_units[i];
    if (r.status == MIC_ENC_OK) { jobs.push_back(GatherJob{r.out_off, tot, r.frame_len}); tot += r.frame_len; }
  }
  out.assign(u1 - u0, {});
  if (jobs.empty()) return 0;
  int rc;
  if ((rc = e->d_compact.ensure(tot + 64))) return rc;
  if ((rc = e->d_jobs.ensure(jobs.size() * sizeof(GatherJob) + 64))) return rc;
  CUDA_TRY(cudaMemcpyAsync(e->d_jobs.p, jobs.data(), jobs.size() * sizeof(GatherJob), cudaMemcpyHostToDevice, e->stream));
  launch_gather_bytes((const GatherJob*)e->d_jobs.p, (int)jobs.size(), (const uint8_t*)e->d_frames.p, (uint8_t*)e->d_compact.p, e->stream);
  std::vector<uint8_t> host(tot);
  CUDA_TRY(cudaMemcpyAsync(host.data(), e->d_compact.p, tot, cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  size_t j = 0;
  for (int i = u0; i < u1; i++) {
    const MicEncUnit& r = e->h_units[i];
    if (r.status == MIC_ENC_OK) { out[i - u0].assign(host.begin() + jobs[j].dst_off, host.begin() + jobs[j].dst_off + r.frame_len); j++; }
  }
  return 0;
}

WaveletGeom enc_wavelet_geom(unsigned rows, unsigned cols, int levels) {   // collectSubbandOrder (waveletfsecompressu16.go:202-241)
  WaveletGeom G;
  memset(&G, 0, sizeof G);
  G.rows = rows; G.cols = cols; G.levels = levels;
  unsigned nr[10], nc[10];
  nr[0] = rows; nc[0] = cols;
  for (int l = 1; l <= levels; l++) { nr[l] = (nr[l - 1] + 1) / 2; nc[l] = (nc[l - 1] + 1) / 2; }
  unsigned pos = 0;
  int n = 0;
  auto add = [&](unsigned y0, unsigned y1, unsigned x0, unsigned x1) {
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

// CompressMultiFrame (multiframecompress.go:179-224) + WriteMIC2 (multiframe.go:49-92)
int micgpu_mic2_compress(const uint16_t* frames, int width, int height, int nframes, uint16_t max_value, int temporal, uint8_t* out, size_t cap,
                         size_t* out_len) {
  if (!frames || !out || width <= 0 || height <= 0) return fail(MICGPU_E_HEADER, "bad argument");
  if (nframes <= 0) return fail(MICGPU_E_HEADER, "no frames to compress");
  micgpu_encoder* e = default_encoder(current_device());
  if (!e) return MICGPU_E_CUDA;
  std::lock_guard<std::mutex> lk(e->mu);
  const unsigned long long fpx = (unsigned long long)width * height;
  e->units.clear();
  for (int i = 0; i < nframes; i++) {
    if (temporal && i > 0)   // compressResidualFrame: RLE(+length words) with resMax, 2-state then 1-state (multiframecompress.go:146-162)
      enc_add_unit(e, MIC_ENC_RLE, (unsigned long long)nframes * fpx + (unsigned long long)(i - 1) * fpx, (unsigned)fpx, 1, 0xFFFFFFFFu, 2);
    else
      enc_add_unit(e, MIC_ENC_SPATIAL, (unsigned long long)i * fpx, (unsigned)width, (unsigned)height, max_value, 2);
  }
  int rc;
  CUDA_TRY(cudaSetDevice(e->device));
  const unsigned long long src_elems = (unsigned long long)nframes * fpx * (temporal ? 2 : 1);
  if ((rc = e->d_src.ensure(src_elems * 2 + 64))) return rc;
  CUDA_TRY(cudaMemcpyAsync(e->d_src.p, frames, (size_t)nframes * fpx * 2, cudaMemcpyHostToDevice, e->stream));
  if (temporal)
    launch_temporal_residual((const uint16_t*)e->d_src.p, (uint16_t*)e->d_src.p + (size_t)nframes * fpx, fpx, nframes, e->sm_count, e->stream);
  if ((rc = enc_run(e, (const uint16_t*)e->d_src.p))) return rc;
  for (int i = 0; i < nframes; i++)
    if (e->h_units[i].status != MIC_ENC_OK) {
      rc = enc_status_to_rc(e->h_units[i].status);
      return fail(rc, "frame %d: %s", i, micgpu_last_error());
    }
  std::vector<std::vector<uint8_t>> blobs;
  if ((rc = fetch_frames(e, 0, nframes, blobs))) return rc;
  std::vector<uint8_t> o;
  o.insert(o.end(), {'M', 'I', 'C', '2'});
  put32(o, (uint32_t)width); put32(o, (uint32_t)height); put32(o, (uint32_t)nframes);
  o.push_back((uint8_t)(0x01 | (temporal ? 0x02 : 0)));
  o.push_back(0); o.push_back(0); o.push_back(0);
  uint32_t off = 0;
  for (auto& b : blobs) { put32(o, off); put32(o, (uint32_t)b.size()); off += (uint32_t)b.size(); }
  for (auto& b : blobs) o.insert(o.end(), b.begin(), b.end());
  if (o.size() > cap) return fail(MICGPU_E_SIZE, "output buffer too small (%zu needed)", o.size());
  memcpy(out, o.data(), o.size());
  if (out_len) *out_len = o.size();
  return 0;
}

// WaveletV2RLEFSECompressU16 / WaveletV2SIMDRLEFSECompressU16 (waveletfsecompressu16.go:303-371,427-490) for n images of one geometry
int micgpu_wavelet_v2_compress_batch(int n, const uint16_t* const* pixels, int rows, int cols, const uint16_t* max_values, int levels,
                                     uint8_t* const* outs, const size_t* caps, size_t* out_lens, int* status) {
  if (n <= 0) return 0;
  if (!pixels || !outs || rows <= 0 || cols <= 0) return fail(MICGPU_E_HEADER, "bad argument");
  micgpu_encoder* e = default_encoder(current_device());
  if (!e) return MICGPU_E_CUDA;
  std::lock_guard<std::mutex> lk(e->mu);
  if (levels < 1) levels = 1;
  if (levels > 8) levels = 8;
  int applied = levels;
  {
    int r = rows, c = cols;
    for (int l = 0; l < levels; l++) {
      if (r < 2 || c < 2) { applied = l; break; }
      r = (r + 1) / 2; c = (c + 1) / 2;
    }
  }
  const unsigned long long px = (unsigned long long)rows * cols;
  const WaveletGeom G = enc_wavelet_geom((unsigned)rows, (unsigned)cols, applied);
  e->units.clear();
  std::vector<int> uoi(n);
  for (int i = 0; i < n; i++) {
    // V holds one word per coefficient, three per escaped coefficient
    const int u = enc_add_unit(e, MIC_ENC_RLE, (unsigned long long)i * 3 * px, (unsigned)std::min<unsigned long long>(3 * px, 0x7FFFFFF0ull), 1, 0xFFFFFFFEu, 4);
    e->units[u].no_ladder = 1;
    uoi[i] = u;
  }
  int rc;
  CUDA_TRY(cudaSetDevice(e->device));
  if ((rc = e->d_img.ensure((size_t)n * px * 2 + 64))) return rc;
  if ((rc = e->d_wA.ensure((size_t)n * px * 4 + 64))) return rc;
  if ((rc = e->d_wB.ensure((size_t)n * px * 4 + 64))) return rc;
  if ((rc = e->d_src.ensure((size_t)n * 3 * px * 2 + 64))) return rc;
  if ((rc = e->d_stats.ensure((size_t)n * sizeof(int) + 64))) return rc;
  for (int i = 0; i < n; i++)
    CUDA_TRY(cudaMemcpyAsync((uint16_t*)e->d_img.p + (size_t)i * px, pixels[i], (size_t)px * 2, cudaMemcpyHostToDevice, e->stream));
  CUDA_TRY(cudaMemcpyAsync(e->d_stats.p, uoi.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, e->stream));
  launch_wavelet_forward((const uint16_t*)e->d_img.p, (int32_t*)e->d_wA.p, (int32_t*)e->d_wB.p, n, G, e->sm_count, e->stream);
  rc = enc_run(e, (const uint16_t*)e->d_src.p, [&](MicEncUnit* du, cudaStream_t st) {
    launch_wavelet_pack((const int32_t*)e->d_wA.p, (uint16_t*)e->d_src.p, du, (const int*)e->d_stats.p, n, G, st);
  });
  if (rc) return rc;
  std::vector<std::vector<uint8_t>> blobs;
  if ((rc = fetch_frames(e, 0, n, blobs))) return rc;
  int first = 0;
  for (int i = 0; i < n; i++) {
    int st = 0;
    if (e->h_units[i].status != MIC_ENC_OK) st = enc_status_to_rc(e->h_units[i].status);
    else if (11 + blobs[i].size() > caps[i]) st = fail(MICGPU_E_SIZE, "image %d: output buffer too small", i);
    if (!st) {
      uint8_t* o = outs[i];
      const uint32_t r32 = (uint32_t)rows, c32 = (uint32_t)cols;
      for (int k = 0; k < 4; k++) { o[k] = (uint8_t)(r32 >> (8 * k)); o[4 + k] = (uint8_t)(c32 >> (8 * k)); }
      o[8] = (uint8_t)max_values[i]; o[9] = (uint8_t)(max_values[i] >> 8);
      o[10] = (uint8_t)applied;
      memcpy(o + 11, blobs[i].data(), blobs[i].size());
      if (out_lens) out_lens[i] = 11 + blobs[i].size();
    }
    if (status) status[i] = st;
    if (!first && st) first = st;
  }
  return first;
}

// WaveletFSECompressU16 (with_rle = 0, waveletfsecompressu16.go:71-123) and WaveletRLEFSECompressU16 (with_rle = 1, :551-623):
// the V1 layouts -- interleaved in-place lifting, levels clamped to [1, 4], coefficients in raster order, then 4-state FSE
// directly or behind the RLE layer (whose header also carries the length of the coefficient stream)
int micgpu_wavelet_v1_compress(const uint16_t* pixels, int rows, int cols, uint16_t max_value, int levels, int with_rle, uint8_t* out, size_t cap,
                               size_t* out_len) {
  if (!pixels || !out || rows <= 0 || cols <= 0) return fail(MICGPU_E_HEADER, "bad argument");
  micgpu_encoder* e = default_encoder(current_device());
  if (!e) return MICGPU_E_CUDA;
  std::lock_guard<std::mutex> lk(e->mu);
  if (levels < 1) levels = 1;
  if (levels > 4) levels = 4;
  int applied = levels;
  {
    int r = rows, c = cols;
    for (int l = 0; l < levels; l++) {
      if (r < 2 || c < 2) { applied = l; break; }
      r = (r + 1) / 2; c = (c + 1) / 2;
    }
  }
  const unsigned long long px = (unsigned long long)rows * cols;
  if (px > 0x7FFFFFFFull / 3) return fail(MICGPU_E_UNSUPPORTED, "image too large for one coefficient stream");
  WaveletGeom G;                       // raster order = one segment that covers the image
  memset(&G, 0, sizeof G);
  G.rows = (unsigned)rows; G.cols = (unsigned)cols; G.levels = applied; G.nseg = 1;
  G.seg_start[0] = 0; G.seg_y0[0] = 0; G.seg_x0[0] = 0; G.seg_w[0] = (unsigned)cols;
  e->units.clear();
  // V holds one word per coefficient, three per escaped coefficient; rleMaxVal as for WaveletV2 (:595-601)
  const int u = enc_add_unit(e, with_rle ? MIC_ENC_RLE : MIC_ENC_RAW, 0, (unsigned)(3 * px), 1, 0xFFFFFFFEu, 4);
  e->units[u].no_ladder = 1;           // FSECompressU16FourState, no fallback
  int rc;
  CUDA_TRY(cudaSetDevice(e->device));
  if ((rc = e->d_img.ensure((size_t)px * 2 + 64))) return rc;
  if ((rc = e->d_wA.ensure((size_t)px * 4 + 64))) return rc;
  if ((rc = e->d_wB.ensure((size_t)px * 4 + 64))) return rc;
  if ((rc = e->d_src.ensure((size_t)3 * px * 2 + 64))) return rc;
  if ((rc = e->d_stats.ensure(sizeof(int) + 64))) return rc;
  CUDA_TRY(cudaMemcpyAsync(e->d_img.p, pixels, (size_t)px * 2, cudaMemcpyHostToDevice, e->stream));
  CUDA_TRY(cudaMemcpyAsync(e->d_stats.p, &u, sizeof(int), cudaMemcpyHostToDevice, e->stream));
  launch_wavelet_forward_v1((const uint16_t*)e->d_img.p, (int32_t*)e->d_wA.p, (int32_t*)e->d_wB.p, 1, (unsigned)rows, (unsigned)cols, applied,
                            e->sm_count, e->stream);
  rc = enc_run(e, (const uint16_t*)e->d_src.p, [&](MicEncUnit* du, cudaStream_t st) {
    launch_wavelet_pack((const int32_t*)e->d_wA.p, (uint16_t*)e->d_src.p, du, (const int*)e->d_stats.p, 1, G, st);
  });
  if (rc) return rc;
  const MicEncUnit& r = e->h_units[0];
  if (r.status != MIC_ENC_OK) return enc_status_to_rc(r.status);
  std::vector<std::vector<uint8_t>> blobs;
  if ((rc = fetch_frames(e, 0, 1, blobs))) return rc;
  const size_t hdr = with_rle ? 15 : 11;
  if (out_len) *out_len = hdr + blobs[0].size();
  if (hdr + blobs[0].size() > cap) return fail(MICGPU_E_SIZE, "output buffer too small");
  const uint32_t r32 = (uint32_t)rows, c32 = (uint32_t)cols, n32 = r.v_len;   // v_len: words of the coefficient stream
  for (int k = 0; k < 4; k++) { out[k] = (uint8_t)(r32 >> (8 * k)); out[4 + k] = (uint8_t)(c32 >> (8 * k)); }
  out[8] = (uint8_t)max_value; out[9] = (uint8_t)(max_value >> 8);
  out[10] = (uint8_t)applied;
  if (with_rle) for (int k = 0; k < 4; k++) out[11 + k] = (uint8_t)(n32 >> (8 * k));
  memcpy(out + hdr, blobs[0].data(), blobs[0].size());
  return 0;
}

int micgpu_wavelet_v2_compress(const uint16_t* pixels, int rows, int cols, uint16_t max_value, int levels, uint8_t* out, size_t cap, size_t* out_len) {
  return micgpu_wavelet_v2_compress_batch(1, &pixels, rows, cols, &max_value, levels, &out, &cap, out_len, nullptr);
}

}  // extern "C"

namespace {

// compressTileBlob for tiles [t0, t1) of a prepared job list (wsicompress.go:312-421): planes -> statistics -> units -> frames -> blobs
struct TileSpec { unsigned long long img_off; unsigned img_w, img_h, x0, y0; };

int wsi_encode_tiles(micgpu_encoder* e, const std::vector<TileSpec>& tiles, size_t t0, size_t t1, int tile_w, int tile_h, int channels, int bps,
                     int ct, std::vector<std::vector<uint8_t>>& blobs_out) {
  const bool rgb = channels == 3 && bps == 8;
  const int nplanes = rgb ? 3 : 1;
  const unsigned long long tpx = (unsigned long long)tile_w * tile_h;
  const size_t nt = t1 - t0;
  int rc;
  if ((rc = e->d_planes.ensure(nt * nplanes * tpx * 2 + 64))) return rc;
  std::vector<TilePlaneJob> jobs(nt);
  std::vector<unsigned long long> poffs(nt * nplanes);
  for (size_t i = 0; i < nt; i++) {
    const TileSpec& T = tiles[t0 + i];
    jobs[i] = TilePlaneJob{T.img_off, (unsigned long long)i * nplanes * tpx, T.img_w, T.img_h, T.x0, T.y0, (unsigned)tile_w, (unsigned)tile_h,
                           rgb ? (ct ? 0u : 1u) : (bps <= 8 ? 2u : 3u), 0u};
    for (int k = 0; k < nplanes; k++) poffs[i * nplanes + k] = (unsigned long long)(i * nplanes + k) * tpx;
  }
  const size_t jbytes = jobs.size() * sizeof(TilePlaneJob), obytes = poffs.size() * 8, sbytes = poffs.size() * 2 * sizeof(unsigned);
  if ((rc = e->d_jobs.ensure(jbytes + obytes + 64))) return rc;
  if ((rc = e->d_stats.ensure(sbytes + 64))) return rc;
  cudaStream_t st = e->stream;
  CUDA_TRY(cudaMemcpyAsync(e->d_jobs.p, jobs.data(), jbytes, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync((uint8_t*)e->d_jobs.p + jbytes, poffs.data(), obytes, cudaMemcpyHostToDevice, st));
  launch_tile_planes((const TilePlaneJob*)e->d_jobs.p, (int)nt, (const uint8_t*)e->d_img.p, (uint16_t*)e->d_planes.p, st);
  launch_plane_stats((const uint16_t*)e->d_planes.p, (const unsigned long long*)((uint8_t*)e->d_jobs.p + jbytes), tpx, (int)poffs.size(),
                     (unsigned*)e->d_stats.p, st);
  std::vector<unsigned> stats(poffs.size() * 2);
  CUDA_TRY(cudaMemcpyAsync(stats.data(), e->d_stats.p, sbytes, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  // constant planes need their value: it is the plane max (all samples equal)
  e->units.clear();
  std::vector<int> unit_of_plane(poffs.size(), -1);
  for (size_t p = 0; p < poffs.size(); p++) {
    if (!stats[2 * p + 1]) continue;                                  // constant plane: modes 0 / 1
    const unsigned mv = std::max(stats[2 * p], 255u);                 // wsicompress.go:398-400
    unit_of_plane[p] = enc_add_unit(e, MIC_ENC_SPATIAL, poffs[p], (unsigned)tile_w, (unsigned)tile_h, mv, 2);
  }
  if (!e->units.empty() && (rc = enc_run(e, (const uint16_t*)e->d_planes.p))) return rc;
  std::vector<std::vector<uint8_t>> frames;
  if ((rc = fetch_frames(e, 0, (int)e->units.size(), frames))) return rc;
  // raw fallback planes (ErrUseRLE / ErrIncompressible -> mode 3, wsicompress.go:403-414)
  std::vector<GatherJob> raw;
  std::vector<size_t> raw_plane;
  unsigned long long rtot = 0;
  for (size_t p = 0; p < poffs.size(); p++) {
    const int u = unit_of_plane[p];
    if (u < 0) continue;
    const int s = e->h_units[u].status;
    if (s == MIC_ENC_OK) continue;
    if (s != MIC_ENC_INCOMPRESSIBLE && s != MIC_ENC_USE_RLE) return fail(enc_status_to_rc(s), "tile %zu plane %zu: FSE compress failed", t0 + p / nplanes, p % nplanes);
    raw.push_back(GatherJob{poffs[p], rtot, tpx});
    raw_plane.push_back(p);
    rtot += tpx * 2;
  }
  std::vector<uint8_t> rawhost(rtot);
  if (!raw.empty()) {
    if ((rc = e->d_compact.ensure(rtot + 64))) return rc;
    if ((rc = e->d_jobs.ensure(raw.size() * sizeof(GatherJob) + 64))) return rc;
    CUDA_TRY(cudaMemcpyAsync(e->d_jobs.p, raw.data(), raw.size() * sizeof(GatherJob), cudaMemcpyHostToDevice, st));
    launch_plane_raw((const GatherJob*)e->d_jobs.p, (int)raw.size(), (const uint16_t*)e->d_planes.p, (uint8_t*)e->d_compact.p, st);
    CUDA_TRY(cudaMemcpyAsync(rawhost.data(), e->d_compact.p, rtot, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  // assemble tile blobs: [Ylen][Colen][Cglen] + plane blobs (wsicompress.go:350-361); grey: the plane blob alone
  size_t ri = 0;
  for (size_t i = 0; i < nt; i++) {
    std::vector<uint8_t> pb[3];
    for (int k = 0; k < nplanes; k++) {
      const size_t p = i * nplanes + k;
      const int u = unit_of_plane[p];
      if (u < 0) {
        const unsigned val = stats[2 * p];
        if (val == 0) pb[k] = {0};
        else pb[k] = {1, (uint8_t)val, (uint8_t)(val >> 8)};
      } else if (e->h_units[u].status == MIC_ENC_OK) {
        pb[k].push_back(2);
        pb[k].insert(pb[k].end(), frames[u].begin(), frames[u].end());
      } else {
        pb[k].push_back(3);
        pb[k].insert(pb[k].end(), rawhost.begin() + raw[ri].dst_off, rawhost.begin() + raw[ri].dst_off + tpx * 2);
        ri++;
      }
    }
    std::vector<uint8_t> blob;
    if (rgb) {
      for (int k = 0; k < 3; k++) put32(blob, (uint32_t)pb[k].size());
      for (int k = 0; k < 3; k++) blob.insert(blob.end(), pb[k].begin(), pb[k].end());
    } else {
      blob = std::move(pb[0]);
    }
    blobs_out.push_back(std::move(blob));
  }
  return 0;
}

}  // namespace

extern "C" {

// CompressRGB (rgbcompress.go:25-27) == compressRGBTileBlob on the whole image, colour transform on
int micgpu_rgb_compress(const uint8_t* rgb, int width, int height, uint8_t* out, size_t cap, size_t* out_len) {
  if (!rgb || !out || width <= 0 || height <= 0) return fail(MICGPU_E_HEADER, "bad argument");
  micgpu_encoder* e = default_encoder(current_device());
  if (!e) return MICGPU_E_CUDA;
  std::lock_guard<std::mutex> lk(e->mu);
  int rc;
  CUDA_TRY(cudaSetDevice(e->device));
  const size_t bytes = (size_t)width * height * 3;
  if ((rc = e->d_img.ensure(bytes + 64))) return rc;
  CUDA_TRY(cudaMemcpyAsync(e->d_img.p, rgb, bytes, cudaMemcpyHostToDevice, e->stream));
  std::vector<TileSpec> tiles{TileSpec{0, (unsigned)width, (unsigned)height, 0, 0}};
  std::vector<std::vector<uint8_t>> blobs;
  if ((rc = wsi_encode_tiles(e, tiles, 0, 1, width, height, 3, 8, 1, blobs))) return rc;
  if (blobs[0].size() > cap) return fail(MICGPU_E_SIZE, "output buffer too small");
  memcpy(out, blobs[0].data(), blobs[0].size());
  if (out_len) *out_len = blobs[0].size();
  return 0;
}

// CompressWSI (wsicompress.go:27-172) + WriteMIC3 (wsiformat.go:99-165).  pyramid_levels <= 0: autoLevelCount.
int micgpu_wsi_compress(const uint8_t* pixels, int width, int height, int channels, int bits_per_sample, int tile_w, int tile_h,
                        int pyramid_levels, uint8_t* out, size_t cap, size_t* out_len) {
  if (!pixels || !out || width <= 0 || height <= 0) return fail(MICGPU_E_HEADER, "bad argument");
  const bool rgb = channels == 3 && bits_per_sample == 8;
  if (!rgb && channels != 1) return fail(MICGPU_E_UNSUPPORTED, "MIC3: %d channels at %d bits is not supported", channels, bits_per_sample);
  if (tile_w == 0) tile_w = 256;
  if (tile_h == 0) tile_h = 256;
  const int ct = channels == 3 ? 1 : 0;            // WSIOptions.defaults (wsiformat.go:86-97)
  int nlv = pyramid_levels;
  if (nlv <= 0) {                                  // autoLevelCount (wsiformat.go:273-285)
    nlv = 1;
    int w = width, h = height;
    while (w > tile_w || h > tile_h) { w /= 2; h /= 2; nlv++; if (w <= 1 && h <= 1) break; }
  }
  micgpu_encoder* e = default_encoder(current_device());
  if (!e) return MICGPU_E_CUDA;
  std::lock_guard<std::mutex> lk(e->mu);
  int rc;
  CUDA_TRY(cudaSetDevice(e->device));
  const int bpp = bits_per_sample == 16 ? channels * 2 : channels;
  // level geometry + pyramid on the device (Downsample2xRGB / Grey, wsicompress.go:45-73)
  struct Lv { int w, h, tx, ty, first; unsigned long long off; };
  std::vector<Lv> lv;
  unsigned long long itot = 0;
  {
    int w = width, h = height;
    for (int i = 0; i < nlv; i++) {
      if (i > 0) {
        if (lv.back().w / 2 == 0 || lv.back().h / 2 == 0) break;
        w = lv.back().w / 2; h = lv.back().h / 2;
      }
      lv.push_back(Lv{w, h, (w + tile_w - 1) / tile_w, (h + tile_h - 1) / tile_h, 0, itot});
      itot += ((unsigned long long)w * h * bpp + 255) & ~255ull;
    }
  }
  nlv = (int)lv.size();
  if ((rc = e->d_img.ensure(itot + 64))) return rc;
  CUDA_TRY(cudaMemcpyAsync(e->d_img.p, pixels, (size_t)width * height * bpp, cudaMemcpyHostToDevice, e->stream));
  for (int i = 1; i < nlv; i++)
    launch_downsample2x((const uint8_t*)e->d_img.p + lv[i - 1].off, (uint8_t*)e->d_img.p + lv[i].off, (unsigned)lv[i - 1].w, (unsigned)lv[i - 1].h,
                        bits_per_sample == 16 ? 1u : (unsigned)channels, bits_per_sample == 16 ? 2u : 1u, e->sm_count, e->stream);
  std::vector<TileSpec> tiles;
  int total = 0;
  for (int l = 0; l < nlv; l++) {
    lv[l].first = total;
    for (int ty = 0; ty < lv[l].ty; ty++)
      for (int tx = 0; tx < lv[l].tx; tx++) tiles.push_back(TileSpec{lv[l].off, (unsigned)lv[l].w, (unsigned)lv[l].h, (unsigned)(tx * tile_w), (unsigned)(ty * tile_h)});
    total += lv[l].tx * lv[l].ty;
  }
  std::vector<std::vector<uint8_t>> blobs;
  const size_t chunk = std::max<size_t>(1, (size_t)(1ull << 29) / ((size_t)tile_w * tile_h * (rgb ? 3 : 1) * 2 * 24));   // ~0.5 GB of worst-case scratch
  for (size_t t0 = 0; t0 < tiles.size(); t0 += chunk)
    if ((rc = wsi_encode_tiles(e, tiles, t0, std::min(tiles.size(), t0 + chunk), tile_w, tile_h, channels, bits_per_sample, ct, blobs))) return rc;
  std::vector<uint8_t> o;
  o.insert(o.end(), {'M', 'I', 'C', '3'});
  put32(o, 1); put32(o, (uint32_t)width); put32(o, (uint32_t)height); put32(o, (uint32_t)tile_w); put32(o, (uint32_t)tile_h);
  o.push_back((uint8_t)channels); o.push_back((uint8_t)(channels >> 8));
  o.push_back((uint8_t)bits_per_sample);
  o.push_back((uint8_t)(0x01 | (ct ? 0x02 : 0)));
  o.push_back((uint8_t)nlv); o.push_back((uint8_t)(nlv >> 8));
  o.push_back(0); o.push_back(0);
  put64(o, (uint64_t)total);
  put64(o, 0);
  for (int l = 0; l < nlv; l++) { put32(o, (uint32_t)lv[l].w); put32(o, (uint32_t)lv[l].h); put32(o, (uint32_t)lv[l].tx); put32(o, (uint32_t)lv[l].ty); put32(o, (uint32_t)lv[l].first); }
  uint64_t off = 0;
  for (auto& b : blobs) { put64(o, off); put64(o, (uint64_t)b.size()); off += b.size(); }
  for (auto& b : blobs) o.insert(o.end(), b.begin(), b.end());
  if (o.size() > cap) return fail(MICGPU_E_SIZE, "output buffer too small (%zu needed)", o.size());
  memcpy(out, o.data(), o.size());
  if (out_len) *out_len = o.size();
  return 0;
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
  for (auto& dev : g_enc_ctx)
    for (auto& e : dev) { delete e; e = nullptr; }
}

}  // extern "C"

// ---- .mic file wrappers: what cmd/mic-compress writes and cmd/mic-wasm / web/mic-decoder.js read -------
extern "C" {

// magic of a .mic file: 1 MIC1 (single frame), 2 MIC2, 3 MIC3, 4 MICR (RGB), 5 PICS, 0 unknown (cmd/mic-wasm/main.go:68-99)
int micgpu_file_kind(const uint8_t* file, size_t len) {
  if (!file || len < 4) return 0;
  if (!memcmp(file, "MIC1", 4)) return 1;
  if (!memcmp(file, "MIC2", 4)) return 2;
  if (!memcmp(file, "MIC3", 4)) return 3;
  if (!memcmp(file, "MICR", 4)) return 4;
  if (!memcmp(file, "PICS", 4)) return 5;
  return 0;
}

static void put_u32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
static uint32_t get_u32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

// writeMicFile (cmd/mic-compress/main.go:26-58): "MIC1", width, height, pipeline 1 (Delta+RLE+FSE), length, frame.
// nstates picks compressImage / compressImage4State / the 8-state variant (main.go:93-105), ladder included.
int micgpu_mic1_compress(const uint16_t* pixels, int width, int height, uint16_t max_value, int nstates, uint8_t* out, size_t cap, size_t* out_len) {
  if (!pixels || !out || width <= 0 || height <= 0) return fail(MICGPU_E_HEADER, "bad argument");
  if (cap < 20) return fail(MICGPU_E_SIZE, "output buffer too small");
  size_t flen = 0;
  const int rc = micgpu_compress_single_frame(pixels, width, height, max_value, nstates, out + 20, cap - 20, &flen);
  if (rc) return rc;
  if (flen > 0xFFFFFFFFull) return fail(MICGPU_E_UNSUPPORTED, "frame does not fit the MIC1 length field");
  memcpy(out, "MIC1", 4);
  put_u32(out + 4, (uint32_t)width); put_u32(out + 8, (uint32_t)height); put_u32(out + 12, 1u); put_u32(out + 16, (uint32_t)flen);
  if (out_len) *out_len = 20 + flen;
  return 0;
}

// the MIC1 branch of decodeMicFile (cmd/mic-wasm/main.go:97-131)
int micgpu_mic1_decompress(const uint8_t* file, size_t len, uint16_t* pixels_out, size_t cap_px, int* width, int* height) {
  if (!file || !pixels_out) return fail(MICGPU_E_HEADER, "null argument");
  if (len < 4 || memcmp(file, "MIC1", 4) != 0) return fail(MICGPU_E_HEADER, "invalid .mic magic");
  if (len < 20) return fail(MICGPU_E_HEADER, "MIC1 file too small");
  const uint32_t w = get_u32(file + 4), h = get_u32(file + 8), pipeline = get_u32(file + 12), clen = get_u32(file + 16);
  if (pipeline != 1) return fail(MICGPU_E_UNSUPPORTED, "unsupported pipeline type");
  if (w == 0 || h == 0 || w > 0x7FFFFFFFu || h > 0x7FFFFFFFu) return fail(MICGPU_E_HEADER, "MIC1: invalid dimensions");
  if (clen > len - 20) return fail(MICGPU_E_HEADER, "MIC1: compressed data extends beyond file");
  if ((unsigned long long)w * h > cap_px) return fail(MICGPU_E_SIZE, "output buffer too small");
  if (width) *width = (int)w;
  if (height) *height = (int)h;
  return micgpu_decompress_single_frame(file + 20, clen, pixels_out, (int)w, (int)h);
}

// writeMICRFile (cmd/mic-compress/main.go:61-91): "MICR", width, height, CompressRGB blob
int micgpu_micr_compress(const uint8_t* rgb, int width, int height, uint8_t* out, size_t cap, size_t* out_len) {
  if (!rgb || !out || width <= 0 || height <= 0) return fail(MICGPU_E_HEADER, "bad argument");
  if (cap < 12) return fail(MICGPU_E_SIZE, "output buffer too small");
  size_t blen = 0;
  const int rc = micgpu_rgb_compress(rgb, width, height, out + 12, cap - 12, &blen);
  if (rc) return rc;
  memcpy(out, "MICR", 4);
  put_u32(out + 4, (uint32_t)width); put_u32(out + 8, (uint32_t)height);
  if (out_len) *out_len = 12 + blen;
  return 0;
}

int micgpu_micr_decompress(const uint8_t* file, size_t len, uint8_t* rgb_out, size_t cap_bytes, int* width, int* height) {
  if (!file || !rgb_out) return fail(MICGPU_E_HEADER, "null argument");
  if (len < 12 || memcmp(file, "MICR", 4) != 0) return fail(MICGPU_E_HEADER, "invalid MICR magic");
  const uint32_t w = get_u32(file + 4), h = get_u32(file + 8);
  if (w == 0 || h == 0 || w > 0x7FFFFFFFu || h > 0x7FFFFFFFu) return fail(MICGPU_E_HEADER, "MICR: invalid dimensions");
  if ((unsigned long long)w * h * 3 > cap_bytes) return fail(MICGPU_E_SIZE, "output buffer too small");
  if (width) *width = (int)w;
  if (height) *height = (int)h;
  return micgpu_rgb_decompress(file + 12, len - 12, (int)w, (int)h, rgb_out);
}

}  // extern "C"
