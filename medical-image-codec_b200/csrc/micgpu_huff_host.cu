// micgpu_huff_host.cu -- host side of the canonical-Huffman ENCODER (the decoder is a unit kind of the
// decode plan, micgpu_host.cu / k_huff.cu).
//
// CanHuffmanCompressU16.Compress (canhuffmancompressu16.go:52-81): GenerateFrequencies -> OptimizeSymbolCount ->
// AddDelimiterToSymbolList -> GenerateCanHuffmanTable -> WriteTable -> codes.  The histogram and the bit emission run on
// the device; the code construction works on at most 65536 list entries and runs here, between the two.
// Tie order: the reference sorts with sort.Slice (unstable); this encoder uses stable sorts (equal frequencies stay in
// ascending symbol order, the delimiter last), the order the parity tests' CPU restatement uses too; the streams decode
// under the reference's decoder, which reads the order from the header.
#include <algorithm>
#include <mutex>
#include <vector>

#include "host_common.h"
#include "mic_device.cuh"

namespace micgpu {
void launch_huff_hist(const uint16_t* d_sym, unsigned long long n, uint32_t* d_hist, int sm_count, cudaStream_t st);
int huff_enc_chunks(unsigned long long n);
void launch_huff_emit(const uint16_t* d_sym, unsigned long long n, const uint32_t* d_enc, int depth, unsigned long long* d_chunk,
                      unsigned long long* d_total, unsigned long long first_bit, uint32_t* d_out, int phase, cudaStream_t st);
}  // namespace micgpu

using namespace micgpu;
using namespace micgpu_host;

namespace {

struct SymFreq { uint16_t symbol; uint32_t freq; };   // freq doubles as the code length (canhuffmancompressu16.go:17-20)

struct HuffEncCtx {
  std::mutex mu;
  int device = -1, sm_count = 148;
  cudaStream_t stream = nullptr;
  DevBuf d_sym, d_hist, d_enc, d_chunk, d_out;
};
HuffEncCtx g_ctx[64];

int bit_len_u(unsigned v) { int n = 0; while (v) { n++; v >>= 1; } return n; }

// CalculateCodeLengthForGivenSlice (:215-299): ascending stable sort, then Moffat & Katajainen's in-place lengths
uint32_t code_lengths(std::vector<SymFreq>& f) {
  std::stable_sort(f.begin(), f.end(), [](const SymFreq& a, const SymFreq& b) { return a.freq < b.freq; });
  const long count = (long)f.size();
  if (count == 0) return 0;
  if (count == 1) { f[0].freq = 0; return 0; }
  f[0].freq += f[1].freq;
  long root = 0, leaf = 2;
  for (long next = 1; next < count - 1; next++) {
    if (leaf >= count || f[root].freq < f[leaf].freq) { f[next].freq = f[root].freq; f[root].freq = (uint32_t)next; root++; }
    else { f[next].freq = f[leaf].freq; leaf++; }
    if (leaf >= count || (root < next && f[root].freq < f[leaf].freq)) { f[next].freq += f[root].freq; f[root].freq = (uint32_t)next; root++; }
    else { f[next].freq += f[leaf].freq; leaf++; }
  }
  f[count - 2].freq = 0;
  for (long next = count - 3; next >= 0; next--) f[next].freq = f[f[next].freq].freq + 1;
  long avbl = 1, used = 0, next = count - 1;
  uint32_t depth = 0;
  root = count - 2;
  while (avbl > 0) {
    while (root >= 0 && f[root].freq == depth) { used++; root--; }
    while (avbl > used) { f[next].freq = depth; next--; avbl--; }
    avbl = 2 * used; depth++; used = 0;
  }
  return f[0].freq;
}

// MSB-first header bits (bitwriterhuff.go:19-39)
struct BitsBE {
  std::vector<uint8_t> bytes;
  unsigned long long nbits = 0;
  void add(uint32_t v, int n) {
    for (int i = n - 1; i >= 0; i--) {
      if ((nbits & 7) == 0) bytes.push_back(0);
      if ((v >> i) & 1u) bytes.back() |= (uint8_t)(0x80u >> (nbits & 7));
      nbits++;
    }
  }
};

}  // namespace

extern "C" {

// CanHuffmanCompressU16.Init + Compress (canhuffmancompressu16.go:46-81)
int micgpu_huff_compress(const uint16_t* symbols, size_t n, uint8_t* out, size_t cap, size_t* out_len) {
  if (!symbols || !out || !out_len) return fail(MICGPU_E_HEADER, "null argument");
  if (n == 0 || n > 0xFFFFFFFFull) return fail(MICGPU_E_HEADER, "huffman: symbol count must be in [1, 2^32)");
  int have = 0;
  if (cudaGetDeviceCount(&have) != cudaSuccess || have <= 0) return fail(MICGPU_E_CUDA, "no CUDA device available (libmicgpu has no CPU fallback)");
  const int dev = current_device();
  if (dev < 0 || dev >= 64) return fail(MICGPU_E_CUDA, "device %d out of range", dev);
  HuffEncCtx& C = g_ctx[dev];
  std::lock_guard<std::mutex> lk(C.mu);
  CUDA_TRY(cudaSetDevice(dev));
  if (!C.stream) {
    CUDA_TRY(cudaStreamCreateWithFlags(&C.stream, cudaStreamNonBlocking));
    cudaDeviceGetAttribute(&C.sm_count, cudaDevAttrMultiProcessorCount, dev);
    C.device = dev;
  }
  int rc;
  if ((rc = C.d_sym.ensure(n * 2 + 64)) || (rc = C.d_hist.ensure(65536 * 4)) || (rc = C.d_enc.ensure(65536 * 4))) return rc;
  const int nchunks = huff_enc_chunks(n);
  if ((rc = C.d_chunk.ensure(((size_t)nchunks + 2) * 8))) return rc;
  CUDA_TRY(cudaMemcpyAsync(C.d_sym.p, symbols, n * 2, cudaMemcpyHostToDevice, C.stream));
  CUDA_TRY(cudaMemsetAsync(C.d_hist.p, 0, 65536 * 4, C.stream));
  launch_huff_hist((const uint16_t*)C.d_sym.p, n, (uint32_t*)C.d_hist.p, C.sm_count, C.stream);
  std::vector<uint32_t> hist(65536);
  CUDA_TRY(cudaMemcpyAsync(hist.data(), C.d_hist.p, 65536 * 4, cudaMemcpyDeviceToHost, C.stream));
  CUDA_TRY(cudaStreamSynchronize(C.stream));
  // GenerateFrequencies (:139-166)
  unsigned max_value = 0;
  for (unsigned i = 0; i < 65536; i++) if (hist[i]) max_value = i;
  const int depth = bit_len_u(max_value);
  const unsigned delim = (1u << depth) - 1u;
  std::vector<SymFreq> list;
  for (unsigned i = 0; i < (1u << depth); i++)
    if (hist[i] && i != delim) list.push_back({(uint16_t)i, hist[i]});
  auto by_freq_desc = [](const SymFreq& a, const SymFreq& b) { return a.freq > b.freq; };
  std::stable_sort(list.begin(), list.end(), by_freq_desc);
  // OptimizeSymbolCount (:168-186): the longest prefix whose code stays within 14 bits
  size_t lo = 0, hi = list.size();
  while (lo < hi) {
    const size_t mid = (lo + hi + 1) / 2;
    std::vector<SymFreq> tmp(list.begin(), list.begin() + mid);
    if (code_lengths(tmp) <= 14) lo = mid; else hi = mid - 1;
  }
  list.resize(lo);
  // AddDelimiterToSymbolList (:190-206)
  uint32_t selected = 0;
  for (const SymFreq& f : list) selected += f.freq;
  list.push_back({(uint16_t)delim, (uint32_t)n - selected});
  std::stable_sort(list.begin(), list.end(), by_freq_desc);
  // GenerateCanHuffmanTable (:208-213, :305-344)
  const int max_len = (int)code_lengths(list);
  if (depth + max_len > 32) return fail(MICGPU_E_UNSUPPORTED, "huffman: pixel depth %d + code length %d exceed 32 bits", depth, max_len);   // Go panics (:61-63)
  std::vector<uint32_t> per(max_len + 1, 0), start(max_len + 1, 0), codes(list.size());
  for (const SymFreq& f : list) per[f.freq]++;
  int prev = 0;
  uint32_t nprev = 0;
  for (int i = 1; i <= max_len; i++)
    if (per[i]) {
      start[i] = prev ? (start[prev] + nprev) << (i - prev) : 0u;
      prev = i; nprev = per[i];
    }
  for (size_t i = 0; i < list.size(); i++) codes[i] = start[list[i].freq]++;
  uint32_t dcode = 0, dlen = 0;
  for (size_t i = 0; i < list.size(); i++)
    if (list[i].symbol == delim) { dcode = codes[i]; dlen = list[i].freq; break; }
  // WriteTable (:119-137)
  BitsBE hdr;
  hdr.add((uint32_t)n, 32); hdr.add(max_value, 16); hdr.add((uint32_t)max_len, 8); hdr.add((uint32_t)list.size(), 16);
  for (const SymFreq& f : list) hdr.add(f.symbol, depth);
  const int len_bits = bit_len_u((unsigned)max_len);
  for (const SymFreq& f : list) hdr.add(f.freq, len_bits);
  // GenerateAllSymbolTable (:83-106)
  std::vector<uint32_t> enc(65536, dcode | (dlen << 24) | 0x80000000u);
  for (size_t i = 0; i < list.size(); i++)
    if (list[i].symbol != delim) enc[list[i].symbol] = codes[i] | (list[i].freq << 24);
  CUDA_TRY(cudaMemcpyAsync(C.d_enc.p, enc.data(), 65536 * 4, cudaMemcpyHostToDevice, C.stream));
  unsigned long long* d_chunk = (unsigned long long*)C.d_chunk.p;
  unsigned long long* d_total = d_chunk + nchunks;
  launch_huff_emit((const uint16_t*)C.d_sym.p, n, (const uint32_t*)C.d_enc.p, depth, d_chunk, d_total, 0, nullptr, 0, C.stream);
  unsigned long long data_bits = 0;
  CUDA_TRY(cudaMemcpyAsync(&data_bits, d_total, 8, cudaMemcpyDeviceToHost, C.stream));
  CUDA_TRY(cudaStreamSynchronize(C.stream));
  // the codes, then maxCodeLength + pixelDepth zero bits, then zero padding to a byte (:65-80, flushAlign)
  const unsigned long long total_bits = hdr.nbits + data_bits + (unsigned)(max_len + depth);
  const size_t total_bytes = (size_t)((total_bits + 7) / 8);
  *out_len = total_bytes;
  if (total_bytes > cap) return fail(MICGPU_E_SIZE, "output buffer holds %zu bytes, stream needs %zu", cap, total_bytes);
  const size_t words = total_bytes / 4 + 3;
  if ((rc = C.d_out.ensure(words * 4))) return rc;
  CUDA_TRY(cudaMemsetAsync(C.d_out.p, 0, words * 4, C.stream));
  CUDA_TRY(cudaMemcpyAsync(C.d_out.p, hdr.bytes.data(), hdr.bytes.size(), cudaMemcpyHostToDevice, C.stream));
  launch_huff_emit((const uint16_t*)C.d_sym.p, n, (const uint32_t*)C.d_enc.p, depth, d_chunk, d_total, hdr.nbits, (uint32_t*)C.d_out.p, 1, C.stream);
  CUDA_TRY(cudaMemcpyAsync(out, C.d_out.p, total_bytes, cudaMemcpyDeviceToHost, C.stream));
  CUDA_TRY(cudaStreamSynchronize(C.stream));
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// DeltaRleCompressU16.Compress -> CanHuffmanCompressU16 (fseu16_test.go:881-889)
int micgpu_delta_rle_huff_compress(const uint16_t* pixels, int width, int height, uint16_t max_value, uint8_t* out, size_t cap,
                                   size_t* out_len) {
  if (!pixels || !out || !out_len || width <= 0 || height <= 0) return fail(MICGPU_E_HEADER, "bad argument");
  const size_t npx = (size_t)width * height;
  std::vector<uint16_t> sym(3 * npx + 4096);   // the RLE worst case (every pixel escaped, plus run headers)
  size_t ns = 0;
  int rc = micgpu_delta_rle_compress(pixels, width, height, max_value, sym.data(), sym.size(), &ns);
  if (rc) return rc;
  return micgpu_huff_compress(sym.data(), ns, out, cap, out_len);
}

}  // extern "C"
