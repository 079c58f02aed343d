/*
 * mic_oracle_huff.c -- CPU oracle, canonical-Huffman back end (TEST INFRASTRUCTURE ONLY; see mic_oracle.h).
 *
 * Restates CanHuffmanCompressU16 / CanHuffmanDecompressU16 (canhuffmancompressu16.go, canhuffmandecompressu16.go,
 * bitwriterhuff.go, bitreaderhuff.go) and the DeltaRle -> Huffman composition of the reference's tests
 * (fseu16_test.go:822-898, deltarlehuffdecompressu16.go:19-39).  SURVEY.md section 8(f).4.
 *
 * Parity status: the DECODER is determined by the stream alone (the header carries the symbol order and the code
 * lengths; codes are rebuilt canonically), so it is a straight restatement.  The ENCODER's bytes depend on how
 * sort.Slice (Go's unstable pdqsort) orders symbols of EQUAL frequency; this restatement uses stable sorts (ties in
 * ascending symbol order, the delimiter last), so its bytes can differ from Go's where frequencies tie while both
 * decode to the same symbols under either decoder.  "Parity unpinned" for encoder bytes; the reference holds no
 * golden Huffman streams (canhuffmancompressu16_test.go only round-trips).
 *
 * Stream (all fields MSB first, bitwriterhuff.go:19-39): u32 symbol count, u16 maxValue, u8 maxCodeLength,
 * u16 list size, list size x pixelDepth bits of symbols, list size x Len8(maxCodeLength) bits of code lengths, the codes
 * (a symbol outside the list = the delimiter's code followed by the symbol in pixelDepth bits), then
 * maxCodeLength + pixelDepth zero bits and zero padding to a byte (canhuffmancompressu16.go:52-81,119-137).
 */
#include <stdlib.h>
#include <string.h>

#include "mic_oracle.h"

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

static int hlen16(u32 v) { int n = 0; while (v) { n++; v >>= 1; } return n; }   /* bits.Len16 / bits.Len8 */

typedef struct { u16 symbol; u32 freq; } symfreq;   /* canhuffmancompressu16.go:17-20 (freq doubles as code length) */

/* ---- MSB-first bit writer (bitwriterhuff.go; flush32 + flushAlign produce ceil(total bits / 8) bytes) ------------- */
typedef struct { u8 *p; size_t cap; u64 nbits; } hbw;
static int hbw_add(hbw *w, u32 value, int bits) {   /* addBits16 / addBits32: value masked to `bits` bits */
  if (bits == 0) return 0;
  if (bits < 32) value &= (1u << bits) - 1u;
  size_t need = (size_t)((w->nbits + (u64)bits + 7) >> 3);
  if (need > w->cap) {
    size_t nc = w->cap ? w->cap * 2 : 64;
    while (nc < need) nc *= 2;
    u8 *np = (u8 *)realloc(w->p, nc);
    if (!np) return ORC_ERR_INTERNAL;
    memset(np + w->cap, 0, nc - w->cap);
    w->p = np; w->cap = nc;
  }
  for (int i = bits - 1; i >= 0; i--) {
    if ((value >> i) & 1u) w->p[w->nbits >> 3] |= (u8)(0x80u >> (w->nbits & 7));
    w->nbits++;
  }
  return 0;
}

/* ---- sorts: stable (merge) by frequency ----------------------------------------------------------------------------- */
static void sort_freq(symfreq *a, size_t n, int descending) {
  if (n < 2) return;
  symfreq *t = (symfreq *)malloc(n * sizeof *t);
  for (size_t w = 1; w < n; w *= 2) {
    for (size_t lo = 0; lo < n; lo += 2 * w) {
      size_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n, i = lo, j = mid, k = lo;
      while (i < mid && j < hi) {
        int take_right = descending ? (a[j].freq > a[i].freq) : (a[j].freq < a[i].freq);
        t[k++] = take_right ? a[j++] : a[i++];
      }
      while (i < mid) t[k++] = a[i++];
      while (j < hi) t[k++] = a[j++];
    }
    memcpy(a, t, n * sizeof *t);
  }
  free(t);
}

/* CalculateCodeLengthForGivenSlice (canhuffmancompressu16.go:215-299): ascending sort, then the in-place minimum
 * redundancy code lengths of Moffat & Katajainen; returns the longest length (that of the rarest symbol) */
static u32 code_lengths(symfreq *f, size_t cnt) {
  sort_freq(f, cnt, 0);
  long count = (long)cnt;
  if (count == 0) return 0;
  if (count == 1) { f[0].freq = 0; return 0; }
  f[0].freq += f[1].freq;
  long root = 0, leaf = 2;
  for (long next = 1; next < count - 1; next++) {
    if (leaf >= count || f[root].freq < f[leaf].freq) { f[next].freq = f[root].freq; f[root].freq = (u32)next; root++; }
    else { f[next].freq = f[leaf].freq; leaf++; }
    if (leaf >= count || (root < next && f[root].freq < f[leaf].freq)) { f[next].freq += f[root].freq; f[root].freq = (u32)next; root++; }
    else { f[next].freq += f[leaf].freq; leaf++; }
  }
  f[count - 2].freq = 0;
  for (long next = count - 3; next >= 0; next--) f[next].freq = f[f[next].freq].freq + 1;
  long avbl = 1, used = 0, next = count - 1;
  u32 dpth = 0;
  root = count - 2;
  while (avbl > 0) {
    while (root >= 0 && f[root].freq == dpth) { used++; root--; }
    while (avbl > used) { f[next].freq = dpth; next--; avbl--; }
    avbl = 2 * used; dpth++; used = 0;
  }
  return f[0].freq;
}

/* CalculateSymbolsPerCodeLength + CalculateSymbolStartForCodeLength + ConstructCanHuffmanTable
 * (canhuffmancompressu16.go:305-344): codes[i] for list entry i (freq = its code length).  -1: a length above max_len
 * (Go indexes out of range and panics) */
static int canonical_codes(const symfreq *list, size_t n, int max_len, u32 *codes) {
  u32 per[64], start[64];
  memset(per, 0, sizeof per);
  memset(start, 0, sizeof start);
  for (size_t i = 0; i < n; i++) {
    if (list[i].freq > (u32)max_len) return -1;
    per[list[i].freq]++;
  }
  int prev = 0;
  u32 nprev = 0;
  for (int i = 1; i <= max_len; i++) {
    u32 ns = per[i];
    if (ns != 0) {
      if (prev == 0) start[i] = 0;
      else start[i] = (start[prev] + nprev) << (i - prev);
      prev = i; nprev = ns;
    }
  }
  for (size_t i = 0; i < n; i++) codes[i] = start[list[i].freq]++;
  return 0;
}

int orc_huff_compress(const u16 *in, size_t n, u8 **out, size_t *out_len) {
  if (!out || !out_len || n > 0xFFFFFFFFull) return ORC_ERR_ARG;
  int rc = 0;
  u32 *hist = (u32 *)calloc(65536, sizeof(u32));
  symfreq *list = (symfreq *)malloc(65537 * sizeof *list), *tmp = (symfreq *)malloc(65537 * sizeof *tmp);
  u32 *codes = (u32 *)malloc(65537 * sizeof(u32));
  u32 *enc = (u32 *)malloc(65536 * sizeof(u32));   /* per symbol: code | len << 24 | delimiter << 31 (GenerateAllSymbolTable) */
  hbw w = {0, 0, 0};
  /* GenerateFrequencies (:139-166) */
  u16 max_value = 0;
  for (size_t i = 0; i < n; i++) { hist[in[i]]++; if (in[i] > max_value) max_value = in[i]; }
  const int depth = hlen16(max_value);
  const u32 delim = (1u << depth) - 1u;
  size_t cnt = 0;
  for (u32 i = 0; i < (1u << depth); i++)
    if (hist[i] > 0 && i != delim) { list[cnt].symbol = (u16)i; list[cnt].freq = hist[i]; cnt++; }
  sort_freq(list, cnt, 1);
  /* OptimizeSymbolCount (:168-186): the longest prefix of the list whose code stays within 14 bits */
  size_t lo = 0, hi = cnt;
  while (lo < hi) {
    size_t mid = (lo + hi + 1) / 2;
    memcpy(tmp, list, mid * sizeof *tmp);
    if (code_lengths(tmp, mid) <= 14) lo = mid; else hi = mid - 1;
  }
  cnt = lo;
  /* AddDelimiterToSymbolList (:190-206) */
  u32 selected = 0;
  for (size_t i = 0; i < cnt; i++) selected += list[i].freq;
  list[cnt].symbol = (u16)delim; list[cnt].freq = (u32)n - selected; cnt++;
  sort_freq(list, cnt, 1);
  /* GenerateCanHuffmanTable (:208-213) */
  const int max_len = (int)code_lengths(list, cnt);
  if (depth + max_len > 32 || max_len > 32) { rc = ORC_ERR_ARG; goto done; }   /* Go panics (:61-63) */
  if (canonical_codes(list, cnt, max_len, codes)) { rc = ORC_ERR_INTERNAL; goto done; }
  /* FindIndexOfDelimiter (:108-117) */
  u32 dcode = 0, dlen = 0;
  for (size_t i = 0; i < cnt; i++) if (list[i].symbol == delim) { dcode = codes[i]; dlen = list[i].freq; break; }
  /* WriteTable (:119-137) */
  if ((rc = hbw_add(&w, (u32)n, 32)) || (rc = hbw_add(&w, max_value, 16)) || (rc = hbw_add(&w, (u32)max_len, 8)) ||
      (rc = hbw_add(&w, (u32)cnt, 16))) goto done;
  for (size_t i = 0; i < cnt; i++) if ((rc = hbw_add(&w, list[i].symbol, depth))) goto done;
  const int len_bits = hlen16((u32)max_len);
  for (size_t i = 0; i < cnt; i++) if ((rc = hbw_add(&w, list[i].freq, len_bits))) goto done;
  /* GenerateAllSymbolTable (:83-106) */
  for (u32 i = 0; i < (1u << depth); i++) enc[i] = dcode | (dlen << 24) | 0x80000000u;
  for (size_t i = 0; i < cnt; i++)
    if (list[i].symbol != delim) enc[list[i].symbol] = codes[i] | (list[i].freq << 24);
  /* Compress (:65-80) */
  for (size_t i = 0; i < n; i++) {
    const u32 e = enc[in[i]];
    if ((rc = hbw_add(&w, e & 0xFFFFFFu, (int)((e >> 24) & 0x7F)))) goto done;
    if (e & 0x80000000u) if ((rc = hbw_add(&w, in[i], depth))) goto done;
  }
  for (int k = 0; k < max_len + depth; k++) if ((rc = hbw_add(&w, 0, 1))) goto done;
  *out_len = (size_t)((w.nbits + 7) >> 3);
  *out = w.p;
  w.p = NULL;
done:
  free(hist); free(list); free(tmp); free(codes); free(enc); free(w.p);
  return rc;
}

/* ---- MSB-first reader.  bitreaderhuff.go keeps a 64-bit window refilled 32 bits (or the last bytes) at a time; on a
 * stream that holds every bit it is asked for that is the plain big-endian bit string read here.  A request past the
 * end (Go: a slice panic in fillFastFwd, or stale window bits) is reported as corruption. ---------------------------- */
typedef struct { const u8 *p; u64 nbits, pos; int over; } hbr;
static u32 hbr_get(hbr *r, int n) {
  if (n == 0) return 0;
  if (r->pos + (u64)n > r->nbits) { r->over = 1; r->pos = r->nbits; return 0; }
  u32 v = 0;
  for (int i = 0; i < n; i++) {
    v = (v << 1) | ((r->p[r->pos >> 3] >> (7 - (r->pos & 7))) & 1u);
    r->pos++;
  }
  return v;
}

int orc_huff_decompress(const u8 *in, size_t len, u16 **out, size_t *out_len) {
  if (!in || !out || !out_len) return ORC_ERR_ARG;
  if (len < 9) return ORC_ERR_CORRUPT;   /* the fixed header is 72 bits */
  hbr r = {in, (u64)len * 8, 0, 0};
  /* ReadTable (canhuffmandecompressu16.go:36-79) */
  const u32 n = hbr_get(&r, 32);
  const u32 max_value = hbr_get(&r, 16);
  const int depth = hlen16(max_value);
  const u32 depth_mask = depth ? 0xFFFFFFFFu >> (32 - depth) : 0u;
  const u32 delim = (1u << depth) - 1u;
  const int max_len = (int)hbr_get(&r, 8);
  if (depth + max_len > 32) return ORC_ERR_CORRUPT;   /* DecompressInit panics (:81-86) */
  if (max_len > 24) return ORC_ERR_ARG;               /* oracle limit on the 2^maxCodeLength table */
  const int len_bits = hlen16((u32)max_len);
  const u32 max_mask = max_len ? 0xFFFFFFFFu >> (32 - max_len) : 0u;
  const int both = max_len + depth;
  const u32 both_mask = both ? 0xFFFFFFFFu >> (32 - both) : 0u;
  const u32 cnt = hbr_get(&r, 16);
  if ((u64)cnt * (u64)(depth + len_bits) > r.nbits - r.pos) return ORC_ERR_CORRUPT;
  if (n > 0x7FFFFFFFu) return ORC_ERR_ARG;            /* oracle limit (a zero-length code lets 9 bytes announce 2^32 symbols) */
  symfreq *list = (symfreq *)malloc(((size_t)cnt + 1) * sizeof *list);
  u32 *codes = (u32 *)malloc(((size_t)cnt + 1) * sizeof(u32));
  u32 *table = (u32 *)calloc((size_t)1 << max_len, sizeof(u32));   /* symbol | codeLen << 16 | isDelimiter << 24 */
  u16 *o = (u16 *)malloc(((size_t)n + 1) * sizeof(u16));
  int rc = 0;
  for (u32 i = 0; i < cnt; i++) list[i].symbol = (u16)hbr_get(&r, depth);
  for (u32 i = 0; i < cnt; i++) list[i].freq = hbr_get(&r, len_bits);
  if (r.over) { rc = ORC_ERR_CORRUPT; goto done; }
  if (canonical_codes(list, cnt, max_len, codes)) { rc = ORC_ERR_CORRUPT; goto done; }
  int max_minus_dlen = 0;
  for (u32 j = 0; j < cnt; j++) {
    const int l = (int)list[j].freq;
    const u64 base = (u64)codes[j] << (max_len - l), span = 1ull << (max_len - l);
    if (base + span > (1ull << max_len)) { rc = ORC_ERR_CORRUPT; goto done; }   /* Go: index out of range */
    const u32 e = list[j].symbol | ((u32)l << 16) | (list[j].symbol == delim ? 1u << 24 : 0u);
    for (u64 i = 0; i < span; i++) table[base + i] = e;
    if (list[j].symbol == delim) max_minus_dlen = max_len - l;
  }
  /* Decompress (:88-137): the window holds the next maxCodeLength + pixelDepth bits */
  u32 win = hbr_get(&r, both);
  for (u32 k = 0; k < n; k++) {
    const u32 e = table[(depth < 32 ? win >> depth : 0u) & max_mask];
    u32 sym = e & 0xFFFFu;
    int used = (int)((e >> 16) & 0xFF);
    if (e >> 24) { sym = (win >> max_minus_dlen) & depth_mask; used += depth; }
    o[k] = (u16)sym;
    win = ((used < 32 ? win << used : 0u) & both_mask) | hbr_get(&r, used);
    if (r.over) { rc = ORC_ERR_CORRUPT; goto done; }
  }
  *out = o; *out_len = n; o = NULL;
done:
  free(list); free(codes); free(table); free(o);
  return rc;
}

/* DeltaRleCompressU16.Compress -> CanHuffmanCompressU16 (fseu16_test.go:881-889) */
int orc_delta_rle_huff_compress(const u16 *px, int width, int height, u16 max_value, u8 **out, size_t *out_len) {
  u16 *sym = NULL;
  size_t ns = 0;
  int rc = orc_delta_rle_compress(px, width, height, max_value, &sym, &ns);
  if (rc) return rc;
  rc = orc_huff_compress(sym, ns, out, out_len);
  free(sym);
  return rc;
}

/* DeltaRleHuffDecompressU16.Decompress (deltarlehuffdecompressu16.go:19-39; rlehuffdecompressu16.go:21-51 pulls the RLE
 * symbols out of the Huffman decoder one at a time -- the same values DeltaRleDecompressU16 reads from the array) */
int orc_delta_rle_huff_decompress(const u8 *in, size_t len, int width, int height, u16 *px_out) {
  u16 *sym = NULL;
  size_t ns = 0;
  int rc = orc_huff_decompress(in, len, &sym, &ns);
  if (rc) return rc;
  rc = orc_delta_rle_decompress(sym, ns, width, height, px_out);
  free(sym);
  return rc;
}
