/*
 * mic_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY; see mic_oracle.h).
 *
 * Plain C99 restatement of the Go semantics of pappuks/medical-image-codec.
 * Citations "file.go:a-b" are relative to the reference repository root.
 * Nothing here is shipped in, linked into, or called from libmicgpu.so.
 */
#include "mic_oracle.h"

#include <stdlib.h>
#include <string.h>

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t i32;
typedef int64_t i64;

void orc_free(void *p) { free(p); }

/* ------------------------------------------------------------------------ */
/* growable buffers                                                          */
/* ------------------------------------------------------------------------ */
typedef struct { u8 *p; size_t n, cap; } bbuf;
typedef struct { u16 *p; size_t n, cap; } wbuf;

static void bb_reserve(bbuf *b, size_t extra) {
  if (b->n + extra <= b->cap) return;
  size_t nc = b->cap ? b->cap * 2 : 256;
  while (nc < b->n + extra) nc *= 2;
  b->p = (u8 *)realloc(b->p, nc);
  b->cap = nc;
}
static void bb_push(bbuf *b, u8 v) { bb_reserve(b, 1); b->p[b->n++] = v; }
static void bb_append(bbuf *b, const void *src, size_t n) {
  if (!n) return;
  bb_reserve(b, n);
  memcpy(b->p + b->n, src, n);
  b->n += n;
}
static void bb_u32(bbuf *b, u32 v) { u8 t[4] = {(u8)v, (u8)(v >> 8), (u8)(v >> 16), (u8)(v >> 24)}; bb_append(b, t, 4); }
static void bb_u64(bbuf *b, u64 v) { bb_u32(b, (u32)v); bb_u32(b, (u32)(v >> 32)); }
static void wb_reserve(wbuf *b, size_t extra) {
  if (b->n + extra <= b->cap) return;
  size_t nc = b->cap ? b->cap * 2 : 256;
  while (nc < b->n + extra) nc *= 2;
  b->p = (u16 *)realloc(b->p, nc * sizeof(u16));
  b->cap = nc;
}
static void wb_push(wbuf *b, u16 v) { wb_reserve(b, 1); b->p[b->n++] = v; }
static void wb_append(wbuf *b, const u16 *src, size_t n) {
  if (!n) return;
  wb_reserve(b, n);
  memcpy(b->p + b->n, src, n * sizeof(u16));
  b->n += n;
}
static u32 rd32(const u8 *p) { return (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24); }
static u64 rd64(const u8 *p) { return (u64)rd32(p) | ((u64)rd32(p + 4) << 32); }
static u16 rd16(const u8 *p) { return (u16)(p[0] | (p[1] << 8)); }

/* bits.Len16 / bits.Len32 */
static int len32(u32 v) { int n = 0; while (v) { n++; v >>= 1; } return n; }
static int len16(u16 v) { return len32(v); }
/* highBits (fseu16.go:170-172): uint32(bits.Len32(val) - 1) */
static u32 high_bits(u32 v) { return (u32)(len32(v) - 1); }

/* ------------------------------------------------------------------------ */
/* L0: bit writer (bitwriter.go:13-175)                                      */
/* The Go writer keeps a 64-bit container and flushes whole bytes at          */
/* flush32()/flush() points chosen so the container never overflows; the      */
/* emitted bytes are the pure LSB-first concatenation of the fields, so this  */
/* restatement flushes after every add.                                       */
/* ------------------------------------------------------------------------ */
typedef struct { u64 c; unsigned nbits; bbuf out; } bitw;

static void bw_drain(bitw *b) {
  while (b->nbits >= 8) { bb_push(&b->out, (u8)b->c); b->c >>= 8; b->nbits -= 8; }
}
/* addBits32NC (bitwriter.go:50-53) */
static void bw_add(bitw *b, u32 value, unsigned bits) {
  u32 mask = bits >= 32 ? 0xFFFFFFFFu : ((1u << bits) - 1u);
  b->c |= (u64)(value & mask) << b->nbits;
  b->nbits += bits;
  bw_drain(b);
}
/* close (bitwriter.go:162-168): 1-bit end mark, then flushAlign */
static void bw_close(bitw *b) {
  bw_add(b, 1, 1);
  if (b->nbits) { bb_push(&b->out, (u8)b->c); b->c = 0; b->nbits = 0; }
}

/* ------------------------------------------------------------------------ */
/* L0: reverse bit reader (bitreader.go:18-120), literal restatement         */
/* ------------------------------------------------------------------------ */
typedef struct { const u8 *in; size_t off; u64 value; unsigned bits_read; } bitr;

static void br_fill_fast(bitr *b) { /* bitreader.go:66-77 */
  if (b->bits_read < 32) return;
  u32 low = rd32(b->in + b->off - 4);
  b->value = (b->value << 32) | (u64)low;
  b->bits_read -= 32;
  b->off -= 4;
}
static void br_fill(bitr *b) { /* bitreader.go:80-98 */
  if (b->bits_read < 32) return;
  if (b->off > 4) {
    u32 low = rd32(b->in + b->off - 4);
    b->value = (b->value << 32) | (u64)low;
    b->bits_read -= 32;
    b->off -= 4;
    return;
  }
  while (b->off > 0) {
    b->value = (b->value << 8) | (u64)b->in[b->off - 1];
    b->bits_read -= 8;
    b->off--;
  }
}
static int br_init(bitr *b, const u8 *in, size_t len) { /* bitreader.go:26-47 */
  if (len < 1) return ORC_ERR_CORRUPT;
  b->in = in;
  b->off = len;
  u8 v = in[len - 1];
  if (v == 0) return ORC_ERR_CORRUPT;
  b->bits_read = 64;
  b->value = 0;
  if (len >= 8) {
    b->value = rd64(in + len - 8); /* fillFastStart, bitreader.go:101-106 */
    b->bits_read = 0;
    b->off -= 8;
  } else {
    br_fill(b);
    br_fill(b);
  }
  b->bits_read += 8 - high_bits((u32)v);
  return 0;
}
static u32 br_get_fast(bitr *b, unsigned n) { /* bitreader.go:56-61 */
  u32 v = (u32)((b->value << (b->bits_read & 63)) >> ((64 - n) & 63));
  b->bits_read += n;
  return v;
}
static u32 br_get(bitr *b, unsigned n) { /* bitreader.go:49-54 */
  if (n == 0 || b->bits_read >= 64) return 0;
  return br_get_fast(b, n);
}
static int br_finished(const bitr *b) { return b->bits_read >= 64 && b->off == 0; } /* :109-111 */
static int br_close(const bitr *b) { return b->bits_read > 64 ? ORC_ERR_CORRUPT : 0; } /* :114-120 */

/* ------------------------------------------------------------------------ */
/* L1: FSE scratch (fseu16.go:40-103)                                        */
/* ------------------------------------------------------------------------ */
#define MAX_SYM 65535
#define MIN_TABLELOG 5
#define MAX_TABLELOG 16
#define DEFAULT_TABLELOG 11
#define TABLELOG_ABS_MAX 17

typedef struct { u32 new_state; u16 symbol; u8 nb_bits; } dec_sym;  /* fseu16.go:48-52 */
typedef struct { i32 delta_find_state; u32 delta_nb_bits; } sym_tt; /* fseu16.go:41-44 */

typedef struct {
  u32 *count;       /* [65536] */
  i32 *norm;        /* [65536] */
  size_t src_len;   /* s.br.remain() */
  u32 symbol_len;
  u8 table_log;     /* actualTableLog */
  int zero_bits;
  /* compression tables */
  u16 *table_symbol;
  u32 *state_table;
  sym_tt *symbol_tt;
  /* decompression table */
  dec_sym *dec_table;
  bbuf out;         /* s.Out: ncount header then bitstream */
} scratch;

static scratch *scratch_new(void) {
  scratch *s = (scratch *)calloc(1, sizeof(scratch));
  s->count = (u32 *)calloc(65536, sizeof(u32));
  s->norm = (i32 *)calloc(65536 + 1, sizeof(i32));
  return s;
}
static void scratch_free(scratch *s) {
  if (!s) return;
  free(s->count); free(s->norm); free(s->table_symbol); free(s->state_table);
  free(s->symbol_tt); free(s->dec_table); free(s->out.p); free(s);
}
static u32 table_step(u32 ts) { return (ts >> 1) + (ts >> 3) + 3; } /* fseu16.go:166-168 */

/* countSimple (fsecompressu16.go:438-462) + countSimpleNative (asm_generic.go:15-23) */
static int count_simple(scratch *s, const u16 *in, size_t n) {
  for (size_t i = 0; i < n; i++) s->count[in[i]]++;
  u32 sym_len = 0, m = 0;
  for (u32 j = MAX_SYM + 1; j > 0; j--) {
    u32 c = s->count[j - 1];
    if (c != 0) {
      if (sym_len == 0) sym_len = j;
      if (c > m) m = c;
    }
  }
  s->symbol_len = sym_len;
  return (int)m;
}

/* minTableLog (fsecompressu16.go:465-472) */
static u8 min_table_log(const scratch *s) {
  u32 min_bits_src = high_bits((u32)(s->src_len - 1)) + 1;
  u32 min_bits_symbols = high_bits((u32)(s->symbol_len - 1)) + 2;
  if (min_bits_src < min_bits_symbols) return (u8)min_bits_src;
  return (u8)min_bits_symbols;
}

/* optimalTableLog (fsecompressu16.go:480-518); uint8 arithmetic kept */
static void optimal_table_log(scratch *s) {
  u8 table_log = DEFAULT_TABLELOG; /* s.TableLog default, fseu16.go:128-130 */
  u8 min_bits = min_table_log(s);
  u8 max_bits_src = (u8)((u8)high_bits((u32)(s->src_len - 1)) - 2);
  if (max_bits_src < table_log) table_log = max_bits_src;
  if (min_bits > table_log) table_log = min_bits;
  u32 density = (u32)s->src_len / s->symbol_len;
  if (s->symbol_len > 512 && density > 16 && table_log < 13) table_log = 13;
  else if (density > 64 && s->symbol_len > 256 && table_log < 12) table_log = 12;
  else if (density > 32 && s->symbol_len > 128 && table_log < 12) table_log = 12;
  if (max_bits_src < table_log) table_log = max_bits_src;
  if (table_log < MIN_TABLELOG) table_log = MIN_TABLELOG;
  if (table_log > MAX_TABLELOG) table_log = MAX_TABLELOG;
  s->table_log = table_log;
}

static const u32 rtb_table[8] = {0, 473195, 504333, 520860, 550000, 700000, 750000, 830000};

/* normalizeCount2 (fsecompressu16.go:582-667) */
static int normalize_count2(scratch *s) {
  const i32 NYA = -2;
  u32 distributed = 0;
  u32 total = (u32)s->src_len;
  u8 tl = s->table_log;
  u32 low_threshold = total >> tl;
  u32 low_one = (total * 3) >> (tl + 1);
  for (u32 i = 0; i < s->symbol_len; i++) {
    u32 cnt = s->count[i];
    if (cnt == 0) { s->norm[i] = 0; continue; }
    if (cnt <= low_threshold) { s->norm[i] = -1; distributed++; total -= cnt; continue; }
    if (cnt <= low_one) { s->norm[i] = 1; distributed++; total -= cnt; continue; }
    s->norm[i] = NYA;
  }
  /* More present symbols than table cells: the Go code underflows toDistribute and spins for
   * ~2^32 iterations (inputs far too small for the coder); the oracle reports it instead. */
  if (distributed >= (1u << tl)) return ORC_ERR_INTERNAL;
  u32 to_distribute = (1u << tl) - distributed;
  if ((total / to_distribute) > low_one) {
    low_one = (total * 3) / (to_distribute * 2);
    for (u32 i = 0; i < s->symbol_len; i++) {
      u32 cnt = s->count[i];
      if (s->norm[i] == NYA && cnt <= low_one) { s->norm[i] = 1; distributed++; total -= cnt; }
    }
    if (distributed >= (1u << tl)) return ORC_ERR_INTERNAL;
    to_distribute = (1u << tl) - distributed;
  }
  if (distributed == s->symbol_len + 1) {
    u32 max_v = 0, max_c = 0;
    for (u32 i = 0; i < s->symbol_len; i++)
      if (s->count[i] > max_c) { max_v = i; max_c = s->count[i]; }
    s->norm[max_v] += (i32)to_distribute;
    return 0;
  }
  if (total == 0) {
    for (u32 i = 0; to_distribute > 0; i = (i + 1) % s->symbol_len)
      if (s->norm[i] > 0) { to_distribute--; s->norm[i]++; }
    return 0;
  }
  u64 v_step_log = 62 - (u64)tl;
  u64 mid = ((u64)1 << (v_step_log - 1)) - 1;
  u64 r_step = ((((u64)1 << v_step_log) * (u64)to_distribute) + mid) / (u64)total;
  u64 tmp_total = mid;
  for (u32 i = 0; i < s->symbol_len; i++) {
    if (s->norm[i] == NYA) {
      u64 end = tmp_total + (u64)s->count[i] * r_step;
      u32 s_start = (u32)(tmp_total >> v_step_log);
      u32 s_end = (u32)(end >> v_step_log);
      u32 weight = s_end - s_start;
      if (weight < 1) return ORC_ERR_INTERNAL;
      s->norm[i] = (i32)weight;
      tmp_total = end;
    }
  }
  return 0;
}

/* normalizeCount (fsecompressu16.go:524-578) */
static int normalize_count(scratch *s) {
  u8 tl = s->table_log;
  u64 scale = 62 - (u64)tl;
  u64 step = ((u64)1 << 62) / (u64)s->src_len;
  u64 v_step = (u64)1 << (scale - 20);
  i32 still = (i32)(1 << tl);
  u32 largest = 0;
  i32 largest_p = 0;
  u32 low_threshold = (u32)(s->src_len >> tl);
  for (u32 i = 0; i < s->symbol_len; i++) {
    u32 cnt = s->count[i];
    if (cnt == 0) { s->norm[i] = 0; continue; }
    if (cnt <= low_threshold) {
      s->norm[i] = -1;
      still--;
    } else {
      i32 proba = (i32)(((u64)cnt * step) >> scale);
      if (proba < 8) {
        u64 rest_to_beat = v_step * (u64)rtb_table[proba];
        u64 v = (u64)cnt * step - ((u64)proba << scale);
        if (v > rest_to_beat) proba++;
      }
      if (proba > largest_p) { largest_p = proba; largest = i; }
      s->norm[i] = proba;
      still -= proba;
    }
  }
  if (-still >= (s->norm[largest] >> 1)) return normalize_count2(s);
  s->norm[largest] += still;
  return 0;
}

/* writeCount (fsecompressu16.go:191-289): header goes to s->out (reset first) */
static int write_count(scratch *s) {
  u8 tl = s->table_log;
  i32 table_size = 1 << tl;
  int previous0 = 0;
  u32 charnum = 0;
  size_t max_header = (((size_t)s->symbol_len * (size_t)tl) >> 3) + 3;
  u32 bit_stream = (u32)(tl - MIN_TABLELOG);
  unsigned bit_count = 4;
  i32 remaining = table_size + 1;
  i32 threshold = table_size;
  unsigned nb_bits = (unsigned)tl + 1;
  u8 *out = (u8 *)calloc(max_header + 16, 1);
  size_t outp = 0;
  while (remaining > 1) {
    if (previous0) {
      u32 start = charnum;
      while (s->norm[charnum] == 0) charnum++;
      while (charnum >= start + 24) {
        start += 24;
        bit_stream += (u32)0xFFFF << bit_count;
        out[outp] = (u8)bit_stream; out[outp + 1] = (u8)(bit_stream >> 8);
        outp += 2;
        bit_stream >>= 16;
      }
      while (charnum >= start + 3) {
        start += 3;
        bit_stream += 3u << bit_count;
        bit_count += 2;
      }
      bit_stream += (u32)(charnum - start) << bit_count;
      bit_count += 2;
      if (bit_count > 16) {
        out[outp] = (u8)bit_stream; out[outp + 1] = (u8)(bit_stream >> 8);
        outp += 2;
        bit_stream >>= 16;
        bit_count -= 16;
      }
    }
    i32 count = s->norm[charnum];
    charnum++;
    i32 max = (2 * threshold - 1) - remaining;
    if (count < 0) remaining += count; else remaining -= count;
    count++;
    if (count >= threshold) count += max;
    bit_stream += (u32)count << bit_count;
    bit_count += nb_bits;
    if (count < max) bit_count--;
    previous0 = (count == 1);
    if (remaining < 1) { free(out); return ORC_ERR_INTERNAL; }
    while (remaining < threshold) { nb_bits--; threshold >>= 1; }
    if (bit_count > 16) {
      out[outp] = (u8)bit_stream; out[outp + 1] = (u8)(bit_stream >> 8);
      outp += 2;
      bit_stream >>= 16;
      bit_count -= 16;
    }
  }
  out[outp] = (u8)bit_stream; out[outp + 1] = (u8)(bit_stream >> 8);
  outp += (bit_count + 7) / 8;
  if (charnum > s->symbol_len) { free(out); return ORC_ERR_INTERNAL; }
  s->out.n = 0;
  bb_append(&s->out, out, outp);
  free(out);
  return 0;
}

/* buildCTable (fsecompressu16.go:329-431) */
static int build_ctable(scratch *s) {
  u32 table_size = 1u << s->table_log;
  u32 high_threshold = table_size - 1;
  i32 *cumul = (i32 *)calloc(MAX_SYM + 3, sizeof(i32));
  free(s->table_symbol); free(s->state_table); free(s->symbol_tt);
  s->table_symbol = (u16 *)calloc(table_size, sizeof(u16));
  s->state_table = (u32 *)calloc(table_size, sizeof(u32));
  u32 tt_size = s->symbol_len < 256 ? 256 : s->symbol_len;
  s->symbol_tt = (sym_tt *)calloc(tt_size, sizeof(sym_tt));
  u16 *table_symbol = s->table_symbol;
  cumul[0] = 0;
  for (u32 u = 0; u < s->symbol_len; u++) {
    i32 v = s->norm[u];
    if (v == -1) {
      cumul[u + 1] = cumul[u] + 1;
      table_symbol[high_threshold] = (u16)u;
      high_threshold--;
    } else {
      cumul[u + 1] = cumul[u] + v;
    }
  }
  if ((u32)cumul[s->symbol_len] != table_size) { free(cumul); return ORC_ERR_INTERNAL; }
  cumul[s->symbol_len] = (i32)table_size + 1;
  s->zero_bits = 0;
  {
    u32 step = table_step(table_size);
    u32 mask = table_size - 1;
    u32 position = 0;
    i32 large_limit = 1 << (s->table_log - 1);
    for (u32 ui = 0; ui < s->symbol_len; ui++) {
      i32 v = s->norm[ui];
      if (v > large_limit) s->zero_bits = 1;
      for (i32 k = 0; k < v; k++) {
        table_symbol[position] = (u16)ui;
        position = (position + step) & mask;
        while (position > high_threshold) position = (position + step) & mask;
      }
    }
    if (position != 0) { free(cumul); return ORC_ERR_INTERNAL; }
  }
  for (u32 u = 0; u < table_size; u++) {
    u16 v = table_symbol[u];
    s->state_table[cumul[v]] = table_size + u;
    cumul[v]++;
  }
  {
    i32 total = 0;
    u8 tl = s->table_log;
    u32 tlv = ((u32)tl << 16) - (1u << tl);
    for (u32 i = 0; i < s->symbol_len; i++) {
      i32 v = s->norm[i];
      if (v == 0) continue;
      if (v == -1 || v == 1) {
        s->symbol_tt[i].delta_nb_bits = tlv;
        s->symbol_tt[i].delta_find_state = total - 1;
        total++;
      } else {
        u32 max_bits_out = (u32)tl - high_bits((u32)(v - 1));
        u32 min_state_plus = (u32)v << max_bits_out;
        s->symbol_tt[i].delta_nb_bits = (max_bits_out << 16) - min_state_plus;
        s->symbol_tt[i].delta_find_state = total - v;
        total += v;
      }
    }
    if (total != (i32)table_size) { free(cumul); return ORC_ERR_INTERNAL; }
  }
  free(cumul);
  return 0;
}

/* cStateU16.encode (fsecompressu16.go:95-100) */
static void cstate_encode(bitw *bw, const scratch *s, u32 *state, u16 sym) {
  sym_tt tt = s->symbol_tt[sym];
  u32 nb_bits_out = (*state + tt.delta_nb_bits) >> 16;
  i32 dst_state = (i32)(*state >> (nb_bits_out & 31)) + tt.delta_find_state;
  bw_add(bw, *state, nb_bits_out);
  *state = s->state_table[dst_state];
}

/* compress / compress2State / compress4State / compress8State
 * (fsecompressu16.go:110-187, fse2state.go:122-199, fse4state.go:100-191,
 * fse8state.go:113-226).  All four walk the source from the last symbol to
 * the first, symbol i on state (i mod N), then write the final states
 * N-1 .. 0 (tableLog bits each) so the decoder reads state 0 first. */
static void tans_encode(scratch *s, const u16 *src, size_t n, int nstates) {
  bitw bw;
  memset(&bw, 0, sizeof bw);
  bw.out = s->out; /* append after the ncount header (bw.reset(s.Out)) */
  u32 st[8];
  for (int k = 0; k < nstates; k++) st[k] = 1u << s->table_log; /* cStateU16.init :88-92 */
  for (size_t i = n; i-- > 0;) cstate_encode(&bw, s, &st[i % (size_t)nstates], src[i]);
  for (int k = nstates - 1; k >= 0; k--) bw_add(&bw, st[k], s->table_log);
  bw_close(&bw);
  s->out = bw.out;
}

/* ransEncSymbolU16 + buildRansEncTable (ransu16.go:58-183) */
typedef struct { u32 freq, bias; u8 k0; u32 threshold; } rans_enc;
static int rans_encode(scratch *s, const u16 *src, size_t n) { /* rans8state.go:106-220 */
  rans_enc *tt = (rans_enc *)calloc(s->symbol_len, sizeof(rans_enc));
  u32 cumul = 0;
  for (u32 sym = 0; sym < s->symbol_len; sym++) {
    i32 v = s->norm[sym];
    if (v <= 0) continue;
    u32 freq = (u32)v;
    u8 k0 = (u8)(s->table_log - (u8)high_bits(freq));
    tt[sym].freq = freq; tt[sym].bias = cumul; tt[sym].k0 = k0; tt[sym].threshold = freq << k0;
    cumul += freq;
  }
  for (u32 sym = 0; sym < s->symbol_len; sym++) {
    if (s->norm[sym] != -1) continue;
    u8 k0 = s->table_log;
    tt[sym].freq = 1; tt[sym].bias = cumul; tt[sym].k0 = k0; tt[sym].threshold = 1u << k0;
    cumul++;
  }
  if (cumul != (1u << s->table_log)) { free(tt); return ORC_ERR_INTERNAL; }
  bitw bw;
  memset(&bw, 0, sizeof bw);
  bw.out = s->out;
  u32 table_size = 1u << s->table_log;
  u32 st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (size_t i = n; i-- > 0;) {
    rans_enc e = tt[src[i]];
    u32 x = st[i & 7];
    u32 xl = x + table_size; /* ransEncodeStep, ransu16.go:187-197 */
    u8 k = e.k0;
    if (xl < e.threshold) k--;
    bw_add(&bw, xl, k);
    st[i & 7] = e.bias + ((xl >> k) - e.freq);
  }
  for (int k = 7; k >= 0; k--) bw_add(&bw, st[k], s->table_log);
  bw_close(&bw);
  s->out = bw.out;
  free(tt);
  return 0;
}

/* shared front half of FSECompressU16* (fsecompressu16.go:19-56 etc.) */
static int fse_prepare_tables(scratch *s, const u16 *in, size_t n, int min_len_reject) {
  if (n <= (size_t)min_len_reject) return ORC_ERR_INCOMPRESSIBLE;
  s->src_len = n;
  int max_count = count_simple(s, in, n);
  if ((size_t)max_count == n) return ORC_ERR_USE_RLE;
  if (max_count == 1 || max_count < (int)(n >> 15)) return ORC_ERR_INCOMPRESSIBLE;
  optimal_table_log(s);
  int rc = normalize_count(s);
  if (rc) return rc;
  return write_count(s);
}

int orc_fse_compress(const u16 *in, size_t n, int coder, u8 **out, size_t *out_len) {
  *out = NULL; *out_len = 0;
  int nstates, reject;
  switch (coder) {
    case ORC_FSE1: nstates = 1; reject = 1; break;
    case ORC_FSE2: nstates = 2; reject = 1; break;
    case ORC_FSE4: nstates = 4; reject = 3; break;
    case ORC_FSE8: nstates = 8; reject = 7; break;
    case ORC_RANS8: nstates = 8; reject = 7; break;
    default: return ORC_ERR_ARG;
  }
  scratch *s = scratch_new();
  int rc = fse_prepare_tables(s, in, n, reject);
  if (rc) { scratch_free(s); return rc; }
  if (coder == ORC_RANS8) {
    rc = rans_encode(s, in, n);
  } else {
    rc = build_ctable(s);
    if (!rc) {
      if (n <= 2 && nstates <= 2) rc = ORC_ERR_INTERNAL; /* "src too small" fsecompressu16.go:111, fse2state.go:123 */
      else tans_encode(s, in, n, nstates);
    }
  }
  if (rc) { scratch_free(s); return rc; }
  if (s->out.n >= n * 2) { scratch_free(s); return ORC_ERR_INCOMPRESSIBLE; }
  bbuf o = {0};
  if (coder != ORC_FSE1) {
    bb_push(&o, 0xFF);
    bb_push(&o, coder == ORC_FSE2 ? 0x02 : coder == ORC_FSE4 ? 0x04 : coder == ORC_FSE8 ? 0x84 : 0x08);
    bb_u32(&o, (u32)n);
  }
  bb_append(&o, s->out.p, s->out.n);
  scratch_free(s);
  *out = o.p; *out_len = o.n;
  return 0;
}

int orc_fse_table_info(const u16 *in, size_t n, int *table_log, int *symbol_len, i32 *norm_out) {
  scratch *s = scratch_new();
  int rc = fse_prepare_tables(s, in, n, 1);
  if (!rc) {
    if (table_log) *table_log = s->table_log;
    if (symbol_len) *symbol_len = (int)s->symbol_len;
    if (norm_out) memcpy(norm_out, s->norm, 65536 * sizeof(i32));
  }
  scratch_free(s);
  return rc;
}

/* readNCount (fsedecompressu16.go:48-167); b/off/iend as in byteReader */
static int read_ncount(scratch *s, const u8 *b, size_t blen, size_t *consumed) {
  u32 charnum = 0;
  int previous0 = 0;
  i64 off = 0;
  i64 iend = (i64)blen;
  if (iend < 4) return ORC_ERR_CORRUPT;
#define U32AT(o) rd32(b + (o))
  u32 bit_stream = U32AT(off);
  unsigned nb_bits = (bit_stream & 0xF) + MIN_TABLELOG;
  if (nb_bits > TABLELOG_ABS_MAX) return ORC_ERR_CORRUPT;
  bit_stream >>= 4;
  unsigned bit_count = 4;
  s->table_log = (u8)nb_bits;
  i32 remaining = (i32)((1 << nb_bits) + 1);
  i32 threshold = (i32)(1 << nb_bits);
  i32 got_total = 0;
  nb_bits++;
  while (remaining > 1) {
    if (previous0) {
      u32 n0 = charnum;
      while ((bit_stream & 0xFFFF) == 0xFFFF) {
        n0 += 24;
        if (off < iend - 5) {
          off += 2;
          bit_stream = U32AT(off) >> bit_count;
        } else {
          bit_stream >>= 16;
          bit_count += 16;
        }
      }
      while ((bit_stream & 3) == 3) {
        n0 += 3;
        bit_stream >>= 2;
        bit_count += 2;
      }
      n0 += bit_stream & 3;
      bit_count += 2;
      if (n0 > MAX_SYM) return ORC_ERR_CORRUPT;
      while (charnum < n0) { s->norm[charnum & 0xffff] = 0; charnum++; }
      if (off <= iend - 7 || off + (i64)(bit_count >> 3) <= iend - 4) {
        off += bit_count >> 3;
        bit_count &= 7;
        bit_stream = U32AT(off) >> bit_count;
      } else {
        bit_stream >>= 2;
      }
    }
    i32 max = (2 * threshold - 1) - remaining;
    i32 count;
    if (((i32)bit_stream & (threshold - 1)) < max) {
      count = (i32)bit_stream & (threshold - 1);
      bit_count += nb_bits - 1;
    } else {
      count = (i32)bit_stream & (2 * threshold - 1);
      if (count >= threshold) count -= max;
      bit_count += nb_bits;
    }
    count--;
    if (count < 0) { remaining += count; got_total -= count; }
    else { remaining -= count; got_total += count; }
    if (charnum > MAX_SYM + 1) return ORC_ERR_CORRUPT; /* Go: masked index, then symbolLen check */
    s->norm[charnum & 0xffff] = count;
    charnum++;
    previous0 = (count == 0);
    while (remaining < threshold) { nb_bits--; threshold >>= 1; }
    if (off <= iend - 7 || off + (i64)(bit_count >> 3) <= iend - 4) {
      off += bit_count >> 3;
      bit_count &= 7;
    } else {
      bit_count -= (unsigned)(8 * (iend - 4 - off));
      off = iend - 4;
    }
    if (off < 0 || off + 4 > iend) return ORC_ERR_CORRUPT; /* Go would panic on the slice */
    bit_stream = U32AT(off) >> (bit_count & 31);
  }
#undef U32AT
  s->symbol_len = charnum;
  if (s->symbol_len <= 1) return ORC_ERR_CORRUPT;
  if (s->symbol_len > MAX_SYM + 1) return ORC_ERR_CORRUPT;
  if (remaining != 1) return ORC_ERR_CORRUPT;
  if (bit_count > 32) return ORC_ERR_CORRUPT;
  if (got_total != (1 << s->table_log)) return ORC_ERR_CORRUPT;
  off += (bit_count + 7) >> 3;
  *consumed = (size_t)off;
  return 0;
}

/* buildDtable (fsedecompressu16.go:198-263) */
static int build_dtable(scratch *s) {
  if (s->table_log > MAX_TABLELOG) return ORC_ERR_CORRUPT; /* Go: decTable of 1<<17 is legal; containers never emit it */
  u32 table_size = 1u << s->table_log;
  u32 high_threshold = table_size - 1;
  free(s->dec_table);
  s->dec_table = (dec_sym *)calloc(table_size, sizeof(dec_sym));
  u32 *symbol_next = (u32 *)calloc(s->symbol_len < 256 ? 256 : s->symbol_len, sizeof(u32));
  s->zero_bits = 0;
  {
    i32 large_limit = 1 << (s->table_log - 1);
    for (u32 i = 0; i < s->symbol_len; i++) {
      i32 v = s->norm[i];
      if (v == -1) {
        s->dec_table[high_threshold].symbol = (u16)i;
        high_threshold--;
        symbol_next[i] = 1;
      } else {
        if (v >= large_limit) s->zero_bits = 1;
        symbol_next[i] = (u32)v;
      }
    }
  }
  {
    u32 mask = table_size - 1;
    u32 step = table_step(table_size);
    u32 position = 0;
    for (u32 ss = 0; ss < s->symbol_len; ss++) {
      i32 v = s->norm[ss];
      for (i32 i = 0; i < v; i++) {
        s->dec_table[position].symbol = (u16)ss;
        position = (position + step) & mask;
        while (position > high_threshold) position = (position + step) & mask;
      }
    }
    if (position != 0) { free(symbol_next); return ORC_ERR_CORRUPT; }
  }
  for (u32 u = 0; u < table_size; u++) {
    u16 symbol = s->dec_table[u].symbol;
    u32 next_state = symbol_next[symbol];
    symbol_next[symbol] = next_state + 1;
    u8 n_bits = (u8)(s->table_log - (u8)high_bits(next_state));
    s->dec_table[u].nb_bits = n_bits;
    u32 new_state = (next_state << n_bits) - table_size;
    if (new_state >= table_size) { free(symbol_next); return ORC_ERR_CORRUPT; }
    if (new_state == u && n_bits == 0) { free(symbol_next); return ORC_ERR_CORRUPT; }
    s->dec_table[u].new_state = new_state;
  }
  free(symbol_next);
  return 0;
}

/* buildRansDecTable (ransu16.go:77-135) */
static int build_rans_dtable(scratch *s) {
  if (s->table_log > MAX_TABLELOG) return ORC_ERR_CORRUPT;
  u32 table_size = 1u << s->table_log;
  free(s->dec_table);
  s->dec_table = (dec_sym *)calloc(table_size, sizeof(dec_sym));
  s->zero_bits = 0;
  i32 large_limit = 1 << (s->table_log - 1);
  u32 slot = 0;
  for (u32 sym = 0; sym < s->symbol_len; sym++) {
    i32 v = s->norm[sym];
    if (v <= 0) continue;
    if (v >= large_limit) s->zero_bits = 1;
    u32 freq = (u32)v;
    for (u32 j = 0; j < freq; j++) {
      if (slot >= table_size) return ORC_ERR_CORRUPT;
      u32 x_next = freq + j;
      u8 nb = (u8)(s->table_log - (u8)high_bits(x_next));
      u32 base = (x_next << nb) - table_size;
      if (base >= table_size) return ORC_ERR_CORRUPT;
      s->dec_table[slot].new_state = base;
      s->dec_table[slot].symbol = (u16)sym;
      s->dec_table[slot].nb_bits = nb;
      slot++;
    }
  }
  for (u32 sym = 0; sym < s->symbol_len; sym++) {
    if (s->norm[sym] != -1) continue;
    if (slot >= table_size) return ORC_ERR_CORRUPT;
    s->dec_table[slot].new_state = 0;
    s->dec_table[slot].symbol = (u16)sym;
    s->dec_table[slot].nb_bits = s->table_log;
    slot++;
  }
  if (slot != table_size) return ORC_ERR_CORRUPT;
  return 0;
}

/* decompress (1-state, fsedecompressu16.go:267-377) */
static int decode_1state(scratch *s, const u8 *bits, size_t blen, wbuf *out) {
  bitr br;
  int rc = br_init(&br, bits, blen);
  if (rc) return rc;
  const dec_sym *dt = s->dec_table;
  u32 state = br_get(&br, s->table_log); /* decoderU16.init :386-390 */
  while (br.off >= 8) {
    for (int half = 0; half < 2; half++) {
      br_fill_fast(&br);
      for (int k = 0; k < 2; k++) {
        dec_sym n = dt[state];
        u32 low = s->zero_bits ? br_get(&br, n.nb_bits) : br_get_fast(&br, n.nb_bits);
        state = n.new_state + low;
        wb_push(out, n.symbol);
      }
    }
  }
  for (;;) {
    if (br_finished(&br) && dt[state].nb_bits > 0) { /* decoderU16.finished :403-405 */
      if (state != 0) wb_push(out, dt[state].symbol);
      break;
    }
    br_fill(&br);
    dec_sym n = dt[state];
    u32 low = br_get(&br, n.nb_bits);
    state = n.new_state + low;
    wb_push(out, n.symbol);
    if (out->n > ((size_t)1 << 31)) return ORC_ERR_CORRUPT; /* DecompressLimit */
  }
  return br_close(&br);
}

/* decompress2State / 4State / 8State / ransDecompress8State
 * (fse2state.go:203-308, fse4state.go:195-353, fse8state.go:230-380,
 * rans8state.go:223-412).  Fill points follow the Go code: initial states are
 * read in pairs with fill() between pairs (4-state: A,B | C | D); main loop
 * fillFast before every pair; tail fill() before every symbol. */
static int decode_nstate(scratch *s, const u8 *bits, size_t blen, int nstates, size_t count, wbuf *out) {
  bitr br;
  int rc = br_init(&br, bits, blen);
  if (rc) return rc;
  const dec_sym *dt = s->dec_table;
  u32 st[8];
  if (nstates == 2) {
    st[0] = br_get(&br, s->table_log);
    st[1] = br_get(&br, s->table_log);
  } else if (nstates == 4) {
    st[0] = br_get(&br, s->table_log);
    st[1] = br_get(&br, s->table_log);
    br_fill(&br);
    st[2] = br_get(&br, s->table_log);
    br_fill(&br);
    st[3] = br_get(&br, s->table_log);
  } else {
    for (int k = 0; k < 8; k += 2) {
      if (k) br_fill(&br);
      st[k] = br_get(&br, s->table_log);
      st[k + 1] = br_get(&br, s->table_log);
    }
  }
  size_t remaining = count;
  size_t group = nstates == 2 ? 4 : (size_t)nstates;
  size_t min_off = nstates == 8 ? 16 : 8;
  wb_reserve(out, count < ((size_t)1 << 28) ? count : ((size_t)1 << 28));
  while (br.off >= min_off && remaining >= group) {
    for (size_t g = 0; g < group; g += 2) {
      br_fill_fast(&br);
      int a = (int)(g % (size_t)nstates), b = (int)((g + 1) % (size_t)nstates);
      dec_sym na = dt[st[a]];
      dec_sym nb = dt[st[b]];
      u32 la = s->zero_bits ? br_get(&br, na.nb_bits) : br_get_fast(&br, na.nb_bits);
      u32 lb = s->zero_bits ? br_get(&br, nb.nb_bits) : br_get_fast(&br, nb.nb_bits);
      st[a] = na.new_state + la;
      st[b] = nb.new_state + lb;
      wb_push(out, na.symbol);
      wb_push(out, nb.symbol);
    }
    remaining -= group;
  }
  int k = 0;
  while (remaining > 0) {
    br_fill(&br);
    dec_sym n = dt[st[k]];
    u32 low = br_get(&br, n.nb_bits);
    st[k] = n.new_state + low;
    wb_push(out, n.symbol);
    remaining--;
    k = (k + 1) % nstates;
  }
  return br_close(&br);
}

int orc_fse_decompress_auto(const u8 *in, size_t len, u16 **out, size_t *out_len) {
  *out = NULL; *out_len = 0;
  int nstates = 1, rans = 0;
  /* FSEDecompressU16Auto dispatch order (fse2state.go:102-116) */
  if (len >= 2 && in[0] == 0xFF && in[1] == 0x84) nstates = 8;
  else if (len >= 2 && in[0] == 0xFF && in[1] == 0x08) { nstates = 8; rans = 1; }
  else if (len >= 2 && in[0] == 0xFF && in[1] == 0x04) nstates = 4;
  else if (len >= 2 && in[0] == 0xFF && in[1] == 0x02) nstates = 2;
  size_t count = 0;
  const u8 *b = in;
  size_t blen = len;
  if (nstates > 1) {
    if (len < 6) return ORC_ERR_CORRUPT;
    count = rd32(in + 2);
    b = in + 6;
    blen = len - 6;
  }
  scratch *s = scratch_new();
  size_t consumed = 0;
  int rc = read_ncount(s, b, blen, &consumed);
  if (!rc) rc = rans ? build_rans_dtable(s) : build_dtable(s);
  wbuf o = {0};
  if (!rc) {
    if (nstates == 1) rc = decode_1state(s, b + consumed, blen - consumed, &o);
    else rc = decode_nstate(s, b + consumed, blen - consumed, nstates, count, &o);
  }
  scratch_free(s);
  if (rc) { free(o.p); return rc; }
  if (!o.p) o.p = (u16 *)malloc(2);
  *out = o.p; *out_len = o.n;
  return 0;
}

/* ------------------------------------------------------------------------ */
/* L2: RLE (rlecompressu16.go, rledecompressu16.go)                          */
/* ------------------------------------------------------------------------ */
typedef struct { wbuf out; u16 *b; size_t bc; u16 mid_count; int same; } rle_enc;

static void rle_init(rle_enc *r, u16 max_value) { /* rlecompressu16.go:15-22 */
  memset(r, 0, sizeof *r);
  int depth = len16(max_value);
  r->mid_count = (u16)((1 << (depth - 1)) - 1);
  r->b = (u16 *)malloc(65540 * sizeof(u16)); /* bc < 65536 even when midCount-1 wraps */
  r->bc = 0;
  r->same = 0;
  wb_push(&r->out, max_value);
}
static void rle_encode(rle_enc *r, u16 symbol) { /* rlecompressu16.go:24-70 */
  size_t bc = r->bc;
  if (bc < 2) { r->b[r->bc++] = symbol; return; }
  u16 prev_plus_one = r->b[bc - 2];
  u16 prev = r->b[bc - 1];
  if (prev_plus_one == prev && prev == symbol) {
    if (!r->same && bc > 2) {
      wb_push(&r->out, (u16)(r->mid_count + (u16)(bc - 2)));
      wb_append(&r->out, r->b, bc - 2);
      memmove(r->b, r->b + bc - 2, 2 * sizeof(u16));
      r->bc = 2;
    }
    r->same = 1;
  } else {
    if (r->same && bc > 2) {
      wb_push(&r->out, (u16)bc);
      wb_push(&r->out, r->b[0]);
      r->bc = 0;
    }
    r->same = 0;
  }
  bc = r->bc;
  if ((int)bc >= (int)(u16)(r->mid_count - 1)) {
    if (r->same) {
      wb_push(&r->out, (u16)(bc - 2));
      wb_push(&r->out, r->b[0]);
    } else {
      wb_push(&r->out, (u16)(r->mid_count + (u16)(bc - 2)));
      wb_append(&r->out, r->b, bc - 2);
    }
    memmove(r->b, r->b + bc - 2, 2 * sizeof(u16));
    r->bc = 2;
  }
  r->b[r->bc++] = symbol;
}
static void rle_flush(rle_enc *r) { /* rlecompressu16.go:72-83 */
  size_t bc = r->bc;
  if (bc > 0) {
    if (r->same) {
      wb_push(&r->out, (u16)bc);
      wb_push(&r->out, r->b[0]);
    } else {
      wb_push(&r->out, (u16)(r->mid_count + (u16)bc));
      wb_append(&r->out, r->b, bc);
    }
  }
}

int orc_rle_compress(const u16 *in, size_t n, u16 max_value, u16 **out, size_t *out_len) {
  if (max_value == 0) return ORC_ERR_ARG; /* Go panics on 1<<-1 */
  rle_enc r;
  rle_init(&r, max_value);
  wb_push(&r.out, (u16)(n >> 16)); /* Compress, rlecompressu16.go:85-93 */
  wb_push(&r.out, (u16)n);
  for (size_t i = 0; i < n; i++) rle_encode(&r, in[i]);
  rle_flush(&r);
  free(r.b);
  *out = r.out.p; *out_len = r.out.n;
  return 0;
}

typedef struct { const u16 *in; size_t n, i; u16 mid_count, c, recurring; int overrun; } rle_dec;
static u16 rle_in(rle_dec *r) { /* Go would panic on an out-of-range read */
  if (r->i >= r->n) { r->overrun = 1; return 0; }
  return r->in[r->i++];
}
static void rle_dec_init(rle_dec *r, const u16 *in, size_t n) { /* rledecompressu16.go:21-30 */
  memset(r, 0, sizeof *r);
  r->in = in; r->n = n;
  u16 max_value = n ? in[0] : 0;
  int depth = len16(max_value);
  r->mid_count = (u16)((1 << (depth > 0 ? depth - 1 : 0)) - 1);
  if (depth == 0) r->overrun = 1;
  r->i = 1;
  r->c = 0;
}
static u16 rle_next2(rle_dec *r) { /* DecodeNext2, rledecompressu16.go:59-85 */
  if (r->c > 0 && r->c < r->mid_count) { r->c--; return r->recurring; }
  if (r->c == 0 || r->c == r->mid_count) {
    r->c = rle_in(r);
    if (r->c <= r->mid_count) {
      r->recurring = rle_in(r);
      r->c--;
      return r->recurring;
    }
  }
  u16 o = rle_in(r);
  r->c--;
  return o;
}

int orc_rle_decompress(const u16 *in, size_t n, u16 **out, size_t *out_len) {
  *out = NULL; *out_len = 0;
  if (n < 3) return ORC_ERR_CORRUPT;
  rle_dec r;
  rle_dec_init(&r, in, n);
  u32 outlen = ((u32)rle_in(&r) << 16);
  outlen += rle_in(&r);
  u16 *o = (u16 *)malloc(((size_t)outlen + 1) * sizeof(u16));
  for (u32 i = 0; i < outlen; i++) o[i] = rle_next2(&r);
  if (r.overrun) { free(o); return ORC_ERR_CORRUPT; }
  *out = o; *out_len = outlen;
  return 0;
}

/* ------------------------------------------------------------------------ */
/* L2: Delta + RLE (deltarlecompressu16.go)                                  */
/* ------------------------------------------------------------------------ */
int orc_delta_rle_compress(const u16 *in, int width, int height, u16 max_value, u16 **out, size_t *out_len) {
  *out = NULL; *out_len = 0;
  if (max_value == 0) return ORC_ERR_ARG;
  int depth = len16(max_value);
  u16 thr = (u16)((1 << (depth - 1)) - 1);
  u16 delim = (u16)((1 << depth) - 1);
  rle_enc r;
  rle_init(&r, delim);
  rle_encode(&r, max_value);
  for (int y = 0; y < height; y++) {
    for (int x = 0; x < width; x++) {
      size_t index = (size_t)y * (size_t)width + (size_t)x;
      int div = 0;
      i32 prev = 0;
      if (x > 0) { prev = (i32)in[index - 1]; div++; }
      if (y > 0) { prev += (i32)in[index - (size_t)width]; div++; }
      if (div == 2) prev >>= 1;
      u16 v = in[index];
      i32 diff = (i32)v - prev;
      i32 ad = diff < 0 ? -diff : diff;
      if ((u16)ad >= thr) {
        rle_encode(&r, delim);
        rle_encode(&r, v);
      } else {
        rle_encode(&r, (u16)((i32)thr + diff));
      }
    }
  }
  rle_flush(&r);
  free(r.b);
  *out = r.out.p; *out_len = r.out.n;
  return 0;
}

int orc_delta_rle_decompress(const u16 *in, size_t n, int width, int height, u16 *o) {
  if (n < 2) return ORC_ERR_CORRUPT;
  rle_dec r;
  rle_dec_init(&r, in, n);
  u16 max_value = rle_next2(&r);
  int depth = len16(max_value);
  if (depth == 0) return ORC_ERR_CORRUPT;
  u16 thr = (u16)((1 << (depth - 1)) - 1);
  u16 delim = (u16)((1 << depth) - 1);
  for (int y = 0; y < height; y++) {
    for (int x = 0; x < width; x++) {
      size_t index = (size_t)y * (size_t)width + (size_t)x;
      u16 v = rle_next2(&r);
      if (v == delim) {
        o[index] = rle_next2(&r);
      } else {
        i32 diff = (i32)v - (i32)thr;
        int div = 0;
        i32 prev = 0;
        if (x > 0) { prev = (i32)o[index - 1]; div++; }
        if (y > 0) { prev += (i32)o[index - (size_t)width]; div++; }
        if (div == 2) prev >>= 1;
        o[index] = (u16)(prev + diff);
      }
    }
  }
  return r.overrun ? ORC_ERR_CORRUPT : 0;
}

/* ------------------------------------------------------------------------ */
/* L2: gradient-adaptive Delta + RLE (deltagradrlecompressu16.go, deltagradcompressu16.go:149-166) */
/* ------------------------------------------------------------------------ */
static i32 grad_predict(i32 w, i32 n, i32 nw, i32 ne) { /* deltagradcompressu16.go:149-166, gradShift = 3 (:147) */
  i32 avg = (w + n) >> 1;
  i32 gw = w - nw; if (gw < 0) gw = -gw;
  i32 gn = n - nw; if (gn < 0) gn = -gn;
  i32 g = gw + gn;
  if (g == 0) return avg;
  i32 corr = (ne - nw) >> 3;
  i32 limit = g >> 1;
  if (corr > limit) corr = limit;
  else if (corr < -limit) corr = -limit;
  return avg + corr;
}

/* neighbours of (x, y) in `a`; deltagradrlecompressu16.go:36-52 (encoder) and :97-121 (decoder) use the same cases */
static i32 grad_context(const u16 *a, int width, int x, int y) {
  size_t index = (size_t)y * (size_t)width + (size_t)x;
  if (x == 0 && y == 0) return 0;
  if (y == 0) return (i32)a[index - 1];
  if (x == 0) return (i32)a[index - (size_t)width];
  i32 w = (i32)a[index - 1], n = (i32)a[index - (size_t)width], nw = (i32)a[index - (size_t)width - 1];
  i32 ne = nw;
  if (x + 1 < width) ne = (i32)a[index - (size_t)width + 1];
  return grad_predict(w, n, nw, ne);
}

int orc_grad_delta_rle_compress(const u16 *in, int width, int height, u16 max_value, u16 **out, size_t *out_len) { /* :26-68 */
  *out = NULL; *out_len = 0;
  if (max_value == 0) return ORC_ERR_ARG;
  int depth = len16(max_value);
  u16 thr = (u16)((1 << (depth - 1)) - 1);
  u16 delim = (u16)((1 << depth) - 1);
  rle_enc r;
  rle_init(&r, delim);
  rle_encode(&r, max_value);
  for (int y = 0; y < height; y++) {
    for (int x = 0; x < width; x++) {
      u16 v = in[(size_t)y * (size_t)width + (size_t)x];
      i32 diff = (i32)v - grad_context(in, width, x, y);
      i32 ad = diff < 0 ? -diff : diff;
      if ((u16)ad >= thr) {
        rle_encode(&r, delim);
        rle_encode(&r, v);
      } else {
        rle_encode(&r, (u16)((i32)thr + diff));
      }
    }
  }
  rle_flush(&r);
  free(r.b);
  *out = r.out.p; *out_len = r.out.n;
  return 0;
}

int orc_grad_delta_rle_decompress(const u16 *in, size_t n, int width, int height, u16 *o) { /* :70-133 */
  if (n < 2) return ORC_ERR_CORRUPT;
  rle_dec r;
  rle_dec_init(&r, in, n);
  u16 max_value = rle_next2(&r);
  int depth = len16(max_value);
  if (depth == 0) return ORC_ERR_CORRUPT;
  u16 thr = (u16)((1 << (depth - 1)) - 1);
  u16 delim = (u16)((1 << depth) - 1);
  for (int y = 0; y < height; y++) {
    for (int x = 0; x < width; x++) {
      size_t index = (size_t)y * (size_t)width + (size_t)x;
      u16 v = rle_next2(&r);
      if (v == delim) o[index] = rle_next2(&r);
      else o[index] = (u16)(grad_context(o, width, x, y) + ((i32)v - (i32)thr));
    }
  }
  return r.overrun ? ORC_ERR_CORRUPT : 0;
}

/* ------------------------------------------------------------------------ */
/* L2: ZigZag, temporal, YCoCg-R, pyramid                                    */
/* ------------------------------------------------------------------------ */
u16 orc_zigzag(int16_t x) { return (u16)((i32)((u32)(i32)x << 1) ^ ((i32)x >> 15)); }
int16_t orc_unzigzag(u16 ux) { return (int16_t)((ux >> 1) ^ (u16)(-(i32)(ux & 1))); }

void orc_temporal_encode(const u16 *cur, const u16 *prev, size_t n, u16 *out) {
  if (!prev) { memcpy(out, cur, n * 2); return; }
  for (size_t i = 0; i < n; i++) out[i] = orc_zigzag((int16_t)((i32)cur[i] - (i32)prev[i]));
}
void orc_temporal_decode(const u16 *res, const u16 *prev, size_t n, u16 *out) {
  if (!prev) { memcpy(out, res, n * 2); return; }
  for (size_t i = 0; i < n; i++) out[i] = (u16)((i32)prev[i] + (i32)orc_unzigzag(res[i]));
}

void orc_ycocg_forward(const u8 *rgb, size_t n, u16 *y, u16 *co, u16 *cg) {
  for (size_t i = 0; i < n; i++) {
    int r = rgb[i * 3], g = rgb[i * 3 + 1], b = rgb[i * 3 + 2];
    int co_v = r - b;
    int t = b + (co_v >> 1);
    int cg_v = g - t;
    int y_v = t + (cg_v >> 1);
    y[i] = (u16)y_v;
    co[i] = orc_zigzag((int16_t)co_v);
    cg[i] = orc_zigzag((int16_t)cg_v);
  }
}
void orc_ycocg_inverse(const u16 *y, const u16 *co, const u16 *cg, size_t n, u8 *rgb) {
  for (size_t i = 0; i < n; i++) {
    int y_v = y[i];
    int co_v = orc_unzigzag(co[i]);
    int cg_v = orc_unzigzag(cg[i]);
    int t = y_v - (cg_v >> 1);
    int g = cg_v + t;
    int b = t - (co_v >> 1);
    int r = co_v + b;
    rgb[i * 3] = (u8)r; rgb[i * 3 + 1] = (u8)g; rgb[i * 3 + 2] = (u8)b;
  }
}

int orc_downsample2x_rgb(const u8 *src, int w, int h, u8 *dst, int *nw, int *nh) {
  int new_w = w / 2, new_h = h / 2;
  *nw = new_w; *nh = new_h;
  if (new_w == 0 || new_h == 0) { *nw = 0; *nh = 0; return 0; }
  for (int y = 0; y < new_h; y++)
    for (int x = 0; x < new_w; x++)
      for (int c = 0; c < 3; c++) {
        size_t sy = (size_t)y * 2, sx = (size_t)x * 2;
        int v00 = src[(sy * w + sx) * 3 + c], v10 = src[(sy * w + sx + 1) * 3 + c];
        int v01 = src[((sy + 1) * w + sx) * 3 + c], v11 = src[((sy + 1) * w + sx + 1) * 3 + c];
        dst[((size_t)y * new_w + x) * 3 + c] = (u8)((v00 + v10 + v01 + v11 + 2) / 4);
      }
  return 0;
}
int orc_downsample2x_grey(const u16 *src, int w, int h, u16 *dst, int *nw, int *nh) {
  int new_w = w / 2, new_h = h / 2;
  *nw = new_w; *nh = new_h;
  if (new_w == 0 || new_h == 0) { *nw = 0; *nh = 0; return 0; }
  for (int y = 0; y < new_h; y++)
    for (int x = 0; x < new_w; x++) {
      size_t sy = (size_t)y * 2, sx = (size_t)x * 2;
      u32 v00 = src[sy * w + sx], v10 = src[sy * w + sx + 1];
      u32 v01 = src[(sy + 1) * w + sx], v11 = src[(sy + 1) * w + sx + 1];
      dst[(size_t)y * new_w + x] = (u16)((v00 + v10 + v01 + v11 + 2) / 4);
    }
  return 0;
}

/* ------------------------------------------------------------------------ */
/* L2: 5/3 lifting (waveletu16.go:26-257).  NB Go precedence: a - (b+c)>>1   */
/* parses as a - ((b+c)>>1).                                                 */
/* ------------------------------------------------------------------------ */
void orc_wt53_forward_1d(i32 *data, int offset, int n, int stride) {
  if (n < 2) return;
  int n_half = n / 2;
  for (int i = 0; i < n_half; i++) {
    size_t odd = (size_t)offset + (size_t)(2 * i + 1) * stride;
    size_t left = (size_t)offset + (size_t)(2 * i) * stride;
    size_t right = (2 * i + 2 < n) ? (size_t)offset + (size_t)(2 * i + 2) * stride : left;
    data[odd] = data[odd] - ((data[left] + data[right]) >> 1);
  }
  int n_low = (n + 1) / 2;
  for (int i = 0; i < n_low; i++) {
    size_t even = (size_t)offset + (size_t)(2 * i) * stride;
    i32 d_right;
    if (2 * i + 1 < n) d_right = data[(size_t)offset + (size_t)(2 * i + 1) * stride];
    else d_right = i > 0 ? data[(size_t)offset + (size_t)(2 * i - 1) * stride] : 0;
    i32 d_left = i > 0 ? data[(size_t)offset + (size_t)(2 * i - 1) * stride] : d_right;
    data[even] = data[even] + ((d_left + d_right + 2) >> 2);
  }
}
void orc_wt53_inverse_1d(i32 *data, int offset, int n, int stride) {
  if (n < 2) return;
  int n_half = n / 2, n_low = (n + 1) / 2;
  for (int i = 0; i < n_low; i++) {
    size_t even = (size_t)offset + (size_t)(2 * i) * stride;
    i32 d_right;
    if (2 * i + 1 < n) d_right = data[(size_t)offset + (size_t)(2 * i + 1) * stride];
    else d_right = i > 0 ? data[(size_t)offset + (size_t)(2 * i - 1) * stride] : 0;
    i32 d_left = i > 0 ? data[(size_t)offset + (size_t)(2 * i - 1) * stride] : d_right;
    data[even] = data[even] - ((d_left + d_right + 2) >> 2);
  }
  for (int i = 0; i < n_half; i++) {
    size_t odd = (size_t)offset + (size_t)(2 * i + 1) * stride;
    size_t left = (size_t)offset + (size_t)(2 * i) * stride;
    size_t right = (2 * i + 2 < n) ? (size_t)offset + (size_t)(2 * i + 2) * stride : left;
    data[odd] = data[odd] + ((data[left] + data[right]) >> 1);
  }
}
void orc_wt53_forward_2d(i32 *data, int rows, int cols, int full_cols) { /* waveletu16.go:162-208 */
  int n_col_low = (cols + 1) / 2, n_row_low = (rows + 1) / 2;
  i32 *row_tmp = (i32 *)malloc((size_t)cols * sizeof(i32));
  i32 *col_tmp = (i32 *)malloc((size_t)rows * sizeof(i32));
  for (int y = 0; y < rows; y++) orc_wt53_forward_1d(data, y * full_cols, cols, 1);
  for (int y = 0; y < rows; y++) {
    size_t start = (size_t)y * full_cols;
    memcpy(row_tmp, data + start, (size_t)cols * sizeof(i32));
    for (int i = 0; i < n_col_low; i++) data[start + i] = row_tmp[2 * i];
    for (int i = 0; i < cols / 2; i++) data[start + n_col_low + i] = row_tmp[2 * i + 1];
  }
  for (int x = 0; x < cols; x++) {
    orc_wt53_forward_1d(data, x, rows, full_cols);
    for (int i = 0; i < rows; i++) col_tmp[i] = data[(size_t)i * full_cols + x];
    for (int i = 0; i < n_row_low; i++) data[(size_t)i * full_cols + x] = col_tmp[2 * i];
    for (int i = 0; i < rows / 2; i++) data[(size_t)(n_row_low + i) * full_cols + x] = col_tmp[2 * i + 1];
  }
  free(row_tmp); free(col_tmp);
}
void orc_wt53_inverse_2d(i32 *data, int rows, int cols, int full_cols) { /* waveletu16.go:212-257 */
  int n_col_low = (cols + 1) / 2, n_row_low = (rows + 1) / 2;
  i32 *row_tmp = (i32 *)malloc((size_t)cols * sizeof(i32));
  i32 *col_tmp = (i32 *)malloc((size_t)rows * sizeof(i32));
  for (int x = 0; x < cols; x++) {
    for (int i = 0; i < n_row_low; i++) col_tmp[2 * i] = data[(size_t)i * full_cols + x];
    for (int i = 0; i < rows / 2; i++) col_tmp[2 * i + 1] = data[(size_t)(n_row_low + i) * full_cols + x];
    for (int i = 0; i < rows; i++) data[(size_t)i * full_cols + x] = col_tmp[i];
    orc_wt53_inverse_1d(data, x, rows, full_cols);
  }
  for (int y = 0; y < rows; y++) {
    size_t start = (size_t)y * full_cols;
    memcpy(row_tmp, data + start, (size_t)cols * sizeof(i32));
    for (int i = 0; i < n_col_low; i++) data[start + 2 * i] = row_tmp[i];
    for (int i = 0; i < cols / 2; i++) data[start + 2 * i + 1] = row_tmp[n_col_low + i];
    orc_wt53_inverse_1d(data, (int)start, cols, 1);
  }
  free(row_tmp); free(col_tmp);
}

/* ------------------------------------------------------------------------ */
/* L3: single-unit pipelines (multiframecompress.go:15-175)                  */
/* ------------------------------------------------------------------------ */
static int fse_ladder(const u16 *sym, size_t n, int nstates, u8 **out, size_t *out_len) {
  static const int order[4] = {ORC_FSE8, ORC_FSE4, ORC_FSE2, ORC_FSE1};
  int rc = ORC_ERR_ARG;
  for (int k = 0; k < 4; k++) {
    if (order[k] > nstates) continue;
    rc = orc_fse_compress(sym, n, order[k], out, out_len);
    if (rc == 0) return 0;
  }
  return rc; /* error of the last (1-state) attempt, as Go wraps it */
}

int orc_compress_single_frame(const u16 *px, int width, int height, u16 max_value, int nstates, u8 **out, size_t *out_len) {
  *out = NULL; *out_len = 0;
  if (nstates != 1 && nstates != 2 && nstates != 4 && nstates != 8) return ORC_ERR_ARG;
  u16 *sym; size_t ns;
  int rc = orc_delta_rle_compress(px, width, height, max_value, &sym, &ns);
  if (rc) return rc;
  rc = fse_ladder(sym, ns, nstates, out, out_len);
  free(sym);
  return rc;
}

int orc_decompress_single_frame(const u8 *in, size_t len, int width, int height, u16 *px_out) {
  u16 *sym; size_t ns;
  int rc = orc_fse_decompress_auto(in, len, &sym, &ns);
  if (rc) return rc;
  rc = orc_delta_rle_decompress(sym, ns, width, height, px_out);
  free(sym);
  return rc;
}

/* CompressSingleFrameGrad / DecompressSingleFrameGrad (multiframecompress.go:111-142): two-state first, then one-state */
int orc_compress_single_frame_grad(const u16 *px, int width, int height, u16 max_value, u8 **out, size_t *out_len) {
  *out = NULL; *out_len = 0;
  u16 *sym; size_t ns;
  int rc = orc_grad_delta_rle_compress(px, width, height, max_value, &sym, &ns);
  if (rc) return rc;
  rc = fse_ladder(sym, ns, 2, out, out_len);
  free(sym);
  return rc;
}

int orc_decompress_single_frame_grad(const u8 *in, size_t len, int width, int height, u16 *px_out) {
  u16 *sym; size_t ns;
  int rc = orc_fse_decompress_auto(in, len, &sym, &ns);
  if (rc) return rc;
  rc = orc_grad_delta_rle_decompress(sym, ns, width, height, px_out);
  free(sym);
  return rc;
}

int orc_compress_residual_frame(const u16 *res, size_t n, u16 max_value, u8 **out, size_t *out_len) {
  *out = NULL; *out_len = 0;
  u16 *sym; size_t ns;
  int rc = orc_rle_compress(res, n, max_value, &sym, &ns);
  if (rc) return rc;
  rc = fse_ladder(sym, ns, 2, out, out_len);
  free(sym);
  return rc;
}

int orc_decompress_residual_frame(const u8 *in, size_t len, u16 **out, size_t *out_len) {
  u16 *sym; size_t ns;
  int rc = orc_fse_decompress_auto(in, len, &sym, &ns);
  if (rc) return rc;
  rc = orc_rle_decompress(sym, ns, out, out_len);
  free(sym);
  return rc;
}

/* ------------------------------------------------------------------------ */
/* L3: WaveletV2 (waveletfsecompressu16.go:13-58, 202-421, 538-546)          */
/* ------------------------------------------------------------------------ */
static u16 zz_enc16(i32 v) { return (u16)((v >> 31) ^ (i32)((u32)v << 1)); }
static i32 zz_dec16(u16 v) { u32 u = v; return (i32)((u >> 1) ^ (u32)(-(i32)(u & 1))); }

static void subband_walk(i32 *linear, i32 *data, int rows, int cols, int full_cols, int levels, int scatter) {
  int nr[10], nc[10];
  nr[0] = rows; nc[0] = cols;
  for (int l = 1; l <= levels; l++) { nr[l] = (nr[l - 1] + 1) / 2; nc[l] = (nc[l - 1] + 1) / 2; }
  size_t pos = 0;
#define VISIT(y, x) do { size_t di = (size_t)(y) * full_cols + (x); if (scatter) data[di] = linear[pos]; else linear[pos] = data[di]; pos++; } while (0)
  for (int y = 0; y < nr[levels]; y++) for (int x = 0; x < nc[levels]; x++) VISIT(y, x);
  for (int l = levels; l >= 1; l--) {
    for (int y = 0; y < nr[l]; y++) for (int x = nc[l]; x < nc[l - 1]; x++) VISIT(y, x);
    for (int y = nr[l]; y < nr[l - 1]; y++) for (int x = 0; x < nc[l]; x++) VISIT(y, x);
    for (int y = nr[l]; y < nr[l - 1]; y++) for (int x = nc[l]; x < nc[l - 1]; x++) VISIT(y, x);
  }
#undef VISIT
}

int orc_wavelet_v2_compress(const u16 *px, int rows, int cols, u16 max_value, int levels, u8 **out, size_t *out_len) {
  *out = NULL; *out_len = 0;
  if (levels < 1) levels = 1;
  if (levels > 8) levels = 8;
  size_t n = (size_t)rows * cols;
  i32 *data = (i32 *)malloc(n * sizeof(i32));
  for (size_t i = 0; i < n; i++) data[i] = (i32)px[i];
  int r = rows, c = cols;
  for (int l = 0; l < levels; l++) {
    if (r < 2 || c < 2) { levels = l; break; }
    orc_wt53_forward_2d(data, r, c, cols);
    r = (r + 1) / 2;
    c = (c + 1) / 2;
  }
  i32 *ordered = (i32 *)malloc(n * sizeof(i32));
  subband_walk(ordered, data, rows, cols, cols, levels, 0);
  wbuf enc = {0};
  for (size_t i = 0; i < n; i++) { /* waveletCoeffsToU16 :28-41 */
    i32 v = ordered[i];
    if (v >= -32767 && v <= 32767) wb_push(&enc, zz_enc16(v));
    else { wb_push(&enc, 65535); wb_push(&enc, (u16)((u32)v >> 16)); wb_push(&enc, (u16)(u32)v); }
  }
  free(ordered); free(data);
  u16 zz_max = 0;
  for (size_t i = 0; i < enc.n; i++) if (enc.p[i] > zz_max) zz_max = enc.p[i];
  int depth = len16(zz_max);
  if (depth < 1) depth = 1;
  u16 rle_max = (u16)((1 << depth) - 1);
  u16 *rle; size_t nrle;
  int rc = orc_rle_compress(enc.p, enc.n, rle_max, &rle, &nrle);
  free(enc.p);
  if (rc) return rc;
  u8 *fse; size_t nfse;
  rc = orc_fse_compress(rle, nrle, ORC_FSE4, &fse, &nfse); /* no fallback :353-356 */
  free(rle);
  if (rc) return rc;
  bbuf o = {0};
  bb_u32(&o, (u32)rows); bb_u32(&o, (u32)cols);
  bb_push(&o, (u8)max_value); bb_push(&o, (u8)(max_value >> 8));
  bb_push(&o, (u8)levels);
  bb_append(&o, fse, nfse);
  free(fse);
  *out = o.p; *out_len = o.n;
  return 0;
}

int orc_wavelet_v2_decompress(const u8 *in, size_t len, u16 **px_out, int *rows_o, int *cols_o) {
  *px_out = NULL;
  if (len < 11) return ORC_ERR_CORRUPT;
  int rows = (int)rd32(in), cols = (int)rd32(in + 4);
  int levels = in[10];
  if (len - 11 < 6 || in[11] != 0xFF || in[12] != 0x04) return ORC_ERR_CORRUPT; /* FSEDecompressU16FourState magic */
  u16 *fse; size_t nfse;
  int rc = orc_fse_decompress_auto(in + 11, len - 11, &fse, &nfse);
  if (rc) return rc;
  u16 *enc; size_t nenc;
  rc = orc_rle_decompress(fse, nfse, &enc, &nenc);
  free(fse);
  if (rc) return rc;
  size_t n = (size_t)rows * cols;
  i32 *ordered = (i32 *)calloc(n ? n : 1, sizeof(i32));
  size_t i = 0, k = 0;
  while (i < nenc && k < n) { /* u16ToWaveletCoeffs :45-58 */
    if (enc[i] != 65535) { ordered[k++] = zz_dec16(enc[i]); i++; }
    else {
      if (i + 2 >= nenc) { free(enc); free(ordered); return ORC_ERR_CORRUPT; }
      ordered[k++] = (i32)(((u32)enc[i + 1] << 16) | (u32)enc[i + 2]);
      i += 3;
    }
  }
  free(enc);
  if (k != n) { free(ordered); return ORC_ERR_CORRUPT; } /* Go panics in scatter */
  i32 *data = (i32 *)calloc(n ? n : 1, sizeof(i32));
  subband_walk(ordered, data, rows, cols, cols, levels, 1);
  free(ordered);
  int dr[10], dc[10];
  int r = rows, c = cols;
  for (int l = 0; l < levels && l < 10; l++) { dr[l] = r; dc[l] = c; r = (r + 1) / 2; c = (c + 1) / 2; }
  for (int l = levels - 1; l >= 0; l--) orc_wt53_inverse_2d(data, dr[l], dc[l], cols);
  u16 *px = (u16 *)malloc((n ? n : 1) * sizeof(u16));
  for (size_t j = 0; j < n; j++) px[j] = (u16)data[j];
  free(data);
  *px_out = px; *rows_o = rows; *cols_o = cols;
  return 0;
}

/* ------------------------------------------------------------------------ */
/* L3: V1 wavelet layouts (waveletfsecompressu16.go:71-189, 551-669)         */
/* ------------------------------------------------------------------------ */
/* WaveletFSECompressU16 (with_rle = 0: 11-byte header) and WaveletRLEFSECompressU16 (with_rle = 1: 15-byte header with the
 * coefficient-stream length).  The transform is waveletForward2DRegion (:167-177): every row, then every column, IN PLACE
 * with interleaved low / high coefficients; the next level works on the top-left (r+1)/2 x (c+1)/2 corner of that
 * interleaved buffer (not on an LL band -- the reason WaveletV2 exists); levels are clamped to [1, 4]; coefficients go
 * out in raster order. */
int orc_wavelet_v1_compress(const u16 *px, int rows, int cols, u16 max_value, int levels, int with_rle, u8 **out, size_t *out_len) {
  *out = NULL; *out_len = 0;
  if (levels < 1) levels = 1;
  if (levels > 4) levels = 4;
  size_t n = (size_t)rows * cols;
  i32 *data = (i32 *)malloc((n ? n : 1) * sizeof(i32));
  for (size_t i = 0; i < n; i++) data[i] = (i32)px[i];
  int r = rows, c = cols;
  for (int l = 0; l < levels; l++) {
    if (r < 2 || c < 2) { levels = l; break; }
    for (int y = 0; y < r; y++) orc_wt53_forward_1d(data, y * cols, c, 1);
    for (int x = 0; x < c; x++) orc_wt53_forward_1d(data, x, r, cols);
    r = (r + 1) / 2;
    c = (c + 1) / 2;
  }
  wbuf enc = {0};
  for (size_t i = 0; i < n; i++) {
    i32 v = data[i];
    if (v >= -32767 && v <= 32767) wb_push(&enc, zz_enc16(v));
    else { wb_push(&enc, 65535); wb_push(&enc, (u16)((u32)v >> 16)); wb_push(&enc, (u16)(u32)v); }
  }
  free(data);
  u16 *sym = enc.p; size_t nsym = enc.n;
  size_t enc_len = enc.n;
  int rc = 0;
  if (with_rle) {
    u16 zz_max = 0;
    for (size_t i = 0; i < enc.n; i++) if (enc.p[i] > zz_max) zz_max = enc.p[i];
    int depth = len16(zz_max);
    if (depth < 1) depth = 1;
    rc = orc_rle_compress(enc.p, enc.n, (u16)((1 << depth) - 1), &sym, &nsym);
    free(enc.p);
    if (rc) return rc;
  }
  u8 *fse; size_t nfse;
  rc = orc_fse_compress(sym, nsym, ORC_FSE4, &fse, &nfse);
  free(sym);
  if (rc) return rc;
  bbuf o = {0};
  bb_u32(&o, (u32)rows); bb_u32(&o, (u32)cols);
  bb_push(&o, (u8)max_value); bb_push(&o, (u8)(max_value >> 8));
  bb_push(&o, (u8)levels);
  if (with_rle) bb_u32(&o, (u32)enc_len);
  bb_append(&o, fse, nfse);
  free(fse);
  *out = o.p; *out_len = o.n;
  return 0;
}

/* WaveletFSEDecompressU16 (:124-163) / WaveletRLEFSEDecompressU16 (:624-669) */
int orc_wavelet_v1_decompress(const u8 *in, size_t len, int with_rle, u16 **px_out, int *rows_o, int *cols_o) {
  *px_out = NULL;
  const size_t hdr = with_rle ? 15 : 11;
  if (len < hdr) return ORC_ERR_CORRUPT;
  int rows = (int)rd32(in), cols = (int)rd32(in + 4);
  int levels = in[10];
  if (len - hdr < 6 || in[hdr] != 0xFF || in[hdr + 1] != 0x04) return ORC_ERR_CORRUPT; /* FSEDecompressU16FourState magic */
  u16 *enc; size_t nenc;
  int rc = orc_fse_decompress_auto(in + hdr, len - hdr, &enc, &nenc);
  if (rc) return rc;
  if (with_rle) {
    u16 *e2; size_t n2;
    rc = orc_rle_decompress(enc, nenc, &e2, &n2);
    free(enc);
    if (rc) return rc;
    enc = e2; nenc = n2;
  }
  size_t n = (size_t)rows * cols;
  i32 *data = (i32 *)calloc(n ? n : 1, sizeof(i32));
  size_t i = 0, k = 0;
  while (i < nenc && k < n) { /* u16ToWaveletCoeffs :45-58 */
    if (enc[i] != 65535) { data[k++] = zz_dec16(enc[i]); i++; }
    else {
      if (i + 2 >= nenc) { free(enc); free(data); return ORC_ERR_CORRUPT; } /* Go: index out of range */
      data[k++] = (i32)(((u32)enc[i + 1] << 16) | (u32)enc[i + 2]);
      i += 3;
    }
  }
  free(enc);
  if (k != n) { free(data); return ORC_ERR_CORRUPT; } /* Go: the inverse transform indexes past the short slice */
  int dr[256], dc[256];
  int r = rows, c = cols;
  for (int l = 0; l < levels; l++) { dr[l] = r; dc[l] = c; r = (r + 1) / 2; c = (c + 1) / 2; }
  for (int l = levels - 1; l >= 0; l--) { /* waveletInverse2DRegion :180-189: columns, then rows */
    for (int x = 0; x < dc[l]; x++) orc_wt53_inverse_1d(data, x, dr[l], cols);
    for (int y = 0; y < dr[l]; y++) orc_wt53_inverse_1d(data, y * cols, dc[l], 1);
  }
  u16 *px = (u16 *)malloc((n ? n : 1) * sizeof(u16));
  for (size_t j = 0; j < n; j++) px[j] = (u16)data[j];
  free(data);
  *px_out = px; *rows_o = rows; *cols_o = cols;
  return 0;
}

/* ------------------------------------------------------------------------ */
/* L4: PICS (parallelstrips.go:55-330)                                       */
/* ------------------------------------------------------------------------ */
int orc_pics_compress(const u16 *px, int width, int height, u16 max_value, int num_strips, int nstates, u8 **out, size_t *out_len) {
  *out = NULL; *out_len = 0;
  if (num_strips <= 0) return ORC_ERR_ARG; /* GOMAXPROCS default is a host property; callers pass it */
  if (num_strips > height) num_strips = height;
  if (num_strips < 1) num_strips = 1;
  int strip_h = (height + num_strips - 1) / num_strips;
  int actual = (height + strip_h - 1) / strip_h;
  u8 **blobs = (u8 **)calloc((size_t)actual, sizeof(u8 *));
  size_t *lens = (size_t *)calloc((size_t)actual, sizeof(size_t));
  int rc = 0;
  for (int s = 0; s < actual && !rc; s++) {
    int y0 = s * strip_h, y1 = y0 + strip_h;
    if (y1 > height) y1 = height;
    rc = orc_compress_single_frame(px + (size_t)y0 * width, width, y1 - y0, max_value, nstates, &blobs[s], &lens[s]);
  }
  if (!rc) {
    bbuf o = {0};
    bb_append(&o, "PICS", 4);
    bb_u32(&o, (u32)width); bb_u32(&o, (u32)height); bb_u32(&o, (u32)actual); bb_u32(&o, (u32)strip_h);
    u32 off = 0;
    for (int s = 0; s < actual; s++) { bb_u32(&o, off); bb_u32(&o, (u32)lens[s]); off += (u32)lens[s]; }
    for (int s = 0; s < actual; s++) bb_append(&o, blobs[s], lens[s]);
    *out = o.p; *out_len = o.n;
  }
  for (int s = 0; s < actual; s++) free(blobs[s]);
  free(blobs); free(lens);
  return rc;
}

int orc_pics_decompress(const u8 *in, size_t len, u16 **px_out, int *width, int *height) {
  *px_out = NULL;
  if (len < 20 || memcmp(in, "PICS", 4) != 0) return ORC_ERR_CORRUPT;
  int w = (int)rd32(in + 4), h = (int)rd32(in + 8), ns = (int)rd32(in + 12), sh = (int)rd32(in + 16);
  if (w <= 0 || h <= 0 || ns <= 0 || sh <= 0) return ORC_ERR_CORRUPT;
  size_t header = 20 + (size_t)ns * 8;
  if (len < header) return ORC_ERR_CORRUPT;
  u16 *o = (u16 *)calloc((size_t)w * h, sizeof(u16));
  int rc = 0;
  for (int s = 0; s < ns && !rc; s++) {
    size_t so = rd32(in + 20 + (size_t)s * 8), sl = rd32(in + 24 + (size_t)s * 8);
    size_t start = header + so, end = start + sl;
    if (end > len || start > end) { rc = ORC_ERR_CORRUPT; break; }
    int y0 = s * sh, y1 = y0 + sh;
    if (y1 > h) y1 = h;
    if (y0 >= h) { rc = ORC_ERR_CORRUPT; break; }
    rc = orc_decompress_single_frame(in + start, sl, w, y1 - y0, o + (size_t)y0 * w);
  }
  if (rc) { free(o); return rc; }
  *px_out = o; *width = w; *height = h;
  return 0;
}

/* ------------------------------------------------------------------------ */
/* L4: PICA (parallelstripsadaptive.go:54-289)                               */
/* ------------------------------------------------------------------------ */
/* adaptiveStripBoundaries (:214-289): equal-cost partition of the rows by their summed |vertical delta|.  The sums are
 * integers below 2^53, so the float64 arithmetic of the reference is reproduced exactly by C doubles. */
int orc_pica_boundaries(const u16 *px, int width, int height, int num_strips, int *starts) {
  if (num_strips >= height) {
    for (int i = 0; i < height; i++) starts[i] = i;
    return height;
  }
  if (num_strips == 1) { starts[0] = 0; return 1; }
  double *cum = (double *)calloc((size_t)height + 1, sizeof(double));
  for (int y = 0; y < height; y++) {
    u64 sum = 0;
    if (y >= 1)
      for (int x = 0; x < width; x++) {
        i32 d = (i32)px[(size_t)y * width + x] - (i32)px[(size_t)(y - 1) * width + x];
        sum += (u64)(d < 0 ? -d : d);
      }
    cum[y + 1] = cum[y] + (double)sum;
  }
  double total = cum[height];
  starts[0] = 0;
  if (total == 0) {
    for (int i = 1; i < num_strips; i++) starts[i] = i * height / num_strips;
  } else {
    for (int i = 1; i < num_strips; i++) {
      double target = total * (double)i / (double)num_strips;
      int lo = starts[i - 1] + 1, hi = height;
      while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (cum[mid] < target) lo = mid + 1; else hi = mid;
      }
      if (lo >= height) lo = height - 1;
      starts[i] = lo;
    }
  }
  free(cum);
  return num_strips;
}

int orc_pica_compress(const u16 *px, int width, int height, u16 max_value, int num_strips, u8 **out, size_t *out_len) { /* :54-139 */
  *out = NULL; *out_len = 0;
  if (num_strips <= 0) return ORC_ERR_ARG; /* GOMAXPROCS default is a host property; callers pass it */
  if (num_strips > height) num_strips = height;
  if (num_strips < 1) num_strips = 1;
  int *starts = (int *)calloc((size_t)height + 1, sizeof(int));
  int actual = orc_pica_boundaries(px, width, height, num_strips, starts);
  u8 **blobs = (u8 **)calloc((size_t)actual, sizeof(u8 *));
  size_t *lens = (size_t *)calloc((size_t)actual, sizeof(size_t));
  u32 *flags = (u32 *)calloc((size_t)actual, sizeof(u32));
  int rc = 0;
  for (int s = 0; s < actual && !rc; s++) {
    int y0 = starts[s], y1 = s + 1 < actual ? starts[s + 1] : height;
    const u16 *strip = px + (size_t)y0 * width;
    u8 *ba = NULL, *bg = NULL; size_t la = 0, lg = 0;
    int e1 = orc_compress_single_frame(strip, width, y1 - y0, max_value, 2, &ba, &la);
    int e2 = orc_compress_single_frame_grad(strip, width, y1 - y0, max_value, &bg, &lg);
    if (e2 == 0 && (e1 != 0 || lg <= la)) { blobs[s] = bg; lens[s] = lg; flags[s] = 1; free(ba); }
    else { blobs[s] = ba; lens[s] = la; flags[s] = 0; free(bg); rc = e1; }
  }
  if (!rc) {
    bbuf o = {0};
    bb_append(&o, "PICA", 4);
    bb_u32(&o, (u32)width); bb_u32(&o, (u32)height); bb_u32(&o, (u32)actual);
    u32 off = 0;
    for (int s = 0; s < actual; s++) {
      bb_u32(&o, (u32)starts[s]); bb_u32(&o, off); bb_u32(&o, (u32)lens[s]); bb_u32(&o, flags[s]);
      off += (u32)lens[s];
    }
    for (int s = 0; s < actual; s++) bb_append(&o, blobs[s], lens[s]);
    *out = o.p; *out_len = o.n;
  }
  for (int s = 0; s < actual; s++) free(blobs[s]);
  free(blobs); free(lens); free(flags); free(starts);
  return rc;
}

int orc_pica_decompress(const u8 *in, size_t len, u16 **px_out, int *width, int *height) { /* :143-212 */
  *px_out = NULL;
  if (len < 16 || memcmp(in, "PICA", 4) != 0) return ORC_ERR_CORRUPT;
  int w = (int)rd32(in + 4), h = (int)rd32(in + 8), ns = (int)rd32(in + 12);
  if (ns < 0) return ORC_ERR_CORRUPT;
  size_t header = 16 + (size_t)ns * 16;
  if (len < header) return ORC_ERR_CORRUPT;
  if (w <= 0 || h <= 0 || ns <= 0) return ORC_ERR_CORRUPT;
  u16 *o = (u16 *)calloc((size_t)w * h, sizeof(u16));
  int rc = 0;
  for (int s = 0; s < ns && !rc; s++) {
    const u8 *e = in + 16 + (size_t)s * 16;
    long long y0 = (long long)rd32(e), so = rd32(e + 4), sl = rd32(e + 8);
    u32 fl = rd32(e + 12);
    long long y1 = s + 1 < ns ? (long long)rd32(e + 16) : h;
    size_t start = header + (size_t)so, end = start + (size_t)sl;
    if (end > len || start > end) { rc = ORC_ERR_CORRUPT; break; }
    /* Go would panic on a slice out of range for rows outside the image; the oracle reports corruption */
    if (y1 <= y0 || y0 < 0 || y1 > h) { rc = ORC_ERR_CORRUPT; break; }
    if (fl & 1u) rc = orc_decompress_single_frame_grad(in + start, (size_t)sl, w, (int)(y1 - y0), o + (size_t)y0 * w);
    else rc = orc_decompress_single_frame(in + start, (size_t)sl, w, (int)(y1 - y0), o + (size_t)y0 * w);
  }
  if (rc) { free(o); return rc; }
  *px_out = o; *width = w; *height = h;
  return 0;
}

/* ------------------------------------------------------------------------ */
/* L4: MIC2 (multiframe.go:49-142, multiframecompress.go:179-315)            */
/* ------------------------------------------------------------------------ */
int orc_mic2_compress(const u16 *frames, int width, int height, int nframes, u16 max_value, int temporal, u8 **out, size_t *out_len) {
  *out = NULL; *out_len = 0;
  if (nframes <= 0) return ORC_ERR_ARG;
  size_t fpx = (size_t)width * height;
  u8 **blobs = (u8 **)calloc((size_t)nframes, sizeof(u8 *));
  size_t *lens = (size_t *)calloc((size_t)nframes, sizeof(size_t));
  u16 *res = (u16 *)malloc((fpx ? fpx : 1) * sizeof(u16));
  int rc = 0;
  for (int i = 0; i < nframes && !rc; i++) {
    const u16 *f = frames + (size_t)i * fpx;
    if (temporal && i > 0) {
      orc_temporal_encode(f, f - fpx, fpx, res);
      u16 res_max = 0;
      for (size_t k = 0; k < fpx; k++) if (res[k] > res_max) res_max = res[k];
      rc = orc_compress_residual_frame(res, fpx, res_max, &blobs[i], &lens[i]);
    } else {
      rc = orc_compress_single_frame(f, width, height, max_value, 2, &blobs[i], &lens[i]);
    }
  }
  free(res);
  if (!rc) {
    bbuf o = {0};
    bb_append(&o, "MIC2", 4);
    bb_u32(&o, (u32)width); bb_u32(&o, (u32)height); bb_u32(&o, (u32)nframes);
    bb_push(&o, (u8)(0x01 | (temporal ? 0x02 : 0)));
    bb_push(&o, 0); bb_push(&o, 0); bb_push(&o, 0);
    u32 off = 0;
    for (int i = 0; i < nframes; i++) { bb_u32(&o, off); bb_u32(&o, (u32)lens[i]); off += (u32)lens[i]; }
    for (int i = 0; i < nframes; i++) bb_append(&o, blobs[i], lens[i]);
    *out = o.p; *out_len = o.n;
  }
  for (int i = 0; i < nframes; i++) free(blobs[i]);
  free(blobs); free(lens);
  return rc;
}

static int mic2_header(const u8 *in, size_t len, int *w, int *h, int *n, int *temporal, size_t *data_off) {
  if (len < 20 || memcmp(in, "MIC2", 4) != 0) return ORC_ERR_CORRUPT;
  *w = (int)rd32(in + 4); *h = (int)rd32(in + 8); *n = (int)rd32(in + 12);
  *temporal = (in[16] & 0x02) != 0;
  if (*n < 0) return ORC_ERR_CORRUPT;
  *data_off = 20 + (size_t)*n * 8;
  if (len < *data_off) return ORC_ERR_CORRUPT;
  return 0;
}
static int mic2_frame(const u8 *in, size_t len, size_t data_off, int idx, const u8 **p, size_t *l) {
  size_t o = rd32(in + 20 + (size_t)idx * 8), n = rd32(in + 24 + (size_t)idx * 8);
  if (data_off + o + n > len) return ORC_ERR_CORRUPT;
  *p = in + data_off + o; *l = n;
  return 0;
}
static int mic2_decode_upto(const u8 *in, size_t len, int last, int keep_all, u16 **out, int *w, int *h, int *n, int *temporal) {
  size_t data_off;
  int rc = mic2_header(in, len, w, h, n, temporal, &data_off);
  if (rc) return rc;
  if (last >= *n) return ORC_ERR_ARG;
  size_t fpx = (size_t)*w * *h;
  int first = (*temporal || keep_all) ? 0 : last;
  size_t nkeep = keep_all ? (size_t)(last + 1) : 1;
  u16 *buf = (u16 *)malloc((nkeep * fpx + 1) * sizeof(u16));
  u16 *prev = NULL, *scratch_prev = keep_all ? NULL : (u16 *)malloc((fpx + 1) * sizeof(u16));
  for (int i = first; i <= last; i++) {
    const u8 *p; size_t l;
    rc = mic2_frame(in, len, data_off, i, &p, &l);
    if (rc) break;
    u16 *dst = keep_all ? buf + (size_t)i * fpx : buf;
    if (*temporal && i > 0) {
      u16 *res; size_t nres;
      rc = orc_decompress_residual_frame(p, l, &res, &nres);
      if (rc) break;
      if (nres != fpx) { free(res); rc = ORC_ERR_CORRUPT; break; }
      if (!keep_all) { memcpy(scratch_prev, buf, fpx * 2); prev = scratch_prev; }
      orc_temporal_decode(res, prev, fpx, dst);
      free(res);
    } else {
      rc = orc_decompress_single_frame(p, l, *w, *h, dst);
      if (rc) break;
    }
    prev = dst;
  }
  free(scratch_prev);
  if (rc) { free(buf); return rc; }
  *out = buf;
  return 0;
}
int orc_mic2_decompress(const u8 *in, size_t len, u16 **frames_out, int *width, int *height, int *nframes, int *temporal) {
  *frames_out = NULL;
  size_t data_off;
  int rc = mic2_header(in, len, width, height, nframes, temporal, &data_off);
  if (rc) return rc;
  if (*nframes == 0) { *frames_out = (u16 *)malloc(2); return 0; }
  return mic2_decode_upto(in, len, *nframes - 1, 1, frames_out, width, height, nframes, temporal);
}
int orc_mic2_decompress_frame(const u8 *in, size_t len, int frame_idx, u16 **px_out, int *width, int *height) {
  *px_out = NULL;
  int n, t;
  if (frame_idx < 0) return ORC_ERR_ARG;
  return mic2_decode_upto(in, len, frame_idx, 0, px_out, width, height, &n, &t);
}

/* ------------------------------------------------------------------------ */
/* L4: MIC3 / RGB (wsicompress.go, wsiformat.go, rgbcompress.go)             */
/* ------------------------------------------------------------------------ */
int orc_wsi_plane_compress(const u16 *plane, int width, int height, u8 **out, size_t *out_len) { /* wsicompress.go:373-421 */
  *out = NULL; *out_len = 0;
  size_t n = (size_t)width * height;
  int is_const = 1;
  u16 val = plane[0], max_val = val;
  for (size_t i = 1; i < n; i++) {
    if (plane[i] != val) is_const = 0;
    if (plane[i] > max_val) max_val = plane[i];
  }
  bbuf o = {0};
  if (is_const) {
    if (val == 0) bb_push(&o, 0);
    else { bb_push(&o, 1); bb_push(&o, (u8)val); bb_push(&o, (u8)(val >> 8)); }
    *out = o.p; *out_len = o.n;
    return 0;
  }
  if (max_val < 255) max_val = 255;
  u8 *c; size_t nc;
  int rc = orc_compress_single_frame(plane, width, height, max_val, 2, &c, &nc);
  if (rc == ORC_ERR_USE_RLE || rc == ORC_ERR_INCOMPRESSIBLE) {
    bb_push(&o, 3);
    for (size_t i = 0; i < n; i++) { bb_push(&o, (u8)plane[i]); bb_push(&o, (u8)(plane[i] >> 8)); }
    *out = o.p; *out_len = o.n;
    return 0;
  }
  if (rc) return rc;
  bb_push(&o, 2);
  bb_append(&o, c, nc);
  free(c);
  *out = o.p; *out_len = o.n;
  return 0;
}

int orc_wsi_plane_decompress(const u8 *in, size_t len, int width, int height, u16 *out) { /* wsicompress.go:487-524 */
  size_t n = (size_t)width * height;
  if (len == 0) return ORC_ERR_CORRUPT;
  switch (in[0]) {
    case 0: memset(out, 0, n * 2); return 0;
    case 1: {
      if (len < 3) return ORC_ERR_CORRUPT;
      u16 v = rd16(in + 1);
      for (size_t i = 0; i < n; i++) out[i] = v;
      return 0;
    }
    case 2: return orc_decompress_single_frame(in + 1, len - 1, width, height, out);
    case 3:
      if (len < 1 + n * 2) return ORC_ERR_CORRUPT;
      for (size_t i = 0; i < n; i++) out[i] = rd16(in + 1 + i * 2);
      return 0;
    default: return ORC_ERR_CORRUPT;
  }
}

int orc_rgb_compress(const u8 *rgb, int width, int height, int color_transform, u8 **out, size_t *out_len) { /* :319-364 */
  *out = NULL; *out_len = 0;
  size_t n = (size_t)width * height;
  u16 *pl = (u16 *)malloc(3 * n * sizeof(u16));
  if (color_transform) orc_ycocg_forward(rgb, n, pl, pl + n, pl + 2 * n);
  else for (size_t i = 0; i < n; i++) { pl[i] = rgb[i * 3]; pl[n + i] = rgb[i * 3 + 1]; pl[2 * n + i] = rgb[i * 3 + 2]; }
  u8 *b[3] = {0, 0, 0}; size_t l[3] = {0, 0, 0};
  int rc = 0;
  for (int k = 0; k < 3 && !rc; k++) rc = orc_wsi_plane_compress(pl + (size_t)k * n, width, height, &b[k], &l[k]);
  free(pl);
  if (!rc) {
    bbuf o = {0};
    bb_u32(&o, (u32)l[0]); bb_u32(&o, (u32)l[1]); bb_u32(&o, (u32)l[2]);
    for (int k = 0; k < 3; k++) bb_append(&o, b[k], l[k]);
    *out = o.p; *out_len = o.n;
  }
  for (int k = 0; k < 3; k++) free(b[k]);
  return rc;
}

int orc_rgb_decompress(const u8 *in, size_t len, int width, int height, int color_transform, u8 *rgb_out) { /* :431-474 */
  if (len < 12) return ORC_ERR_CORRUPT;
  size_t l0 = rd32(in), l1 = rd32(in + 4), l2 = rd32(in + 8);
  if (12 + l0 + l1 + l2 > len) return ORC_ERR_CORRUPT;
  size_t n = (size_t)width * height;
  u16 *pl = (u16 *)malloc(3 * n * sizeof(u16) + 2);
  int rc = orc_wsi_plane_decompress(in + 12, l0, width, height, pl);
  if (!rc) rc = orc_wsi_plane_decompress(in + 12 + l0, l1, width, height, pl + n);
  if (!rc) rc = orc_wsi_plane_decompress(in + 12 + l0 + l1, l2, width, height, pl + 2 * n);
  if (!rc) {
    if (color_transform) orc_ycocg_inverse(pl, pl + n, pl + 2 * n, n, rgb_out);
    else for (size_t i = 0; i < n; i++) { rgb_out[i * 3] = (u8)pl[i]; rgb_out[i * 3 + 1] = (u8)pl[n + i]; rgb_out[i * 3 + 2] = (u8)pl[2 * n + i]; }
  }
  free(pl);
  return rc;
}

typedef struct { int w, h, tx, ty, first; } wsi_level;

static int bytes_per_pixel(int channels, int bps) { return bps == 16 ? channels * 2 : channels; }

/* extractTileRGB (wsicompress.go:529-555) */
static void extract_tile(const u8 *img, int iw, int ih, int tw, int th, int tx, int ty, int bpp, u8 *tile) {
  memset(tile, 0, (size_t)tw * th * bpp);
  int sx = tx * tw, sy = ty * th;
  for (int y = 0; y < th; y++) {
    int srcy = sy + y;
    if (srcy >= ih) break;
    int cw = tw;
    if (sx + cw > iw) cw = iw - sx;
    if (cw <= 0) break;
    memcpy(tile + (size_t)y * tw * bpp, img + ((size_t)srcy * iw + sx) * bpp, (size_t)cw * bpp);
  }
}

/* compressTileBlob (wsicompress.go:312-369) */
static int compress_tile_blob(const u8 *tile, int tw, int th, int channels, int bps, int ct, u8 **out, size_t *out_len) {
  if (channels == 3 && bps == 8) return orc_rgb_compress(tile, tw, th, ct, out, out_len);
  size_t n = (size_t)tw * th * channels; /* bytesToUint16Slice :578-590 */
  u16 *pl = (u16 *)malloc((n + 1) * sizeof(u16));
  if (bps <= 8) for (size_t i = 0; i < n; i++) pl[i] = tile[i];
  else for (size_t i = 0; i < n; i++) pl[i] = rd16(tile + i * 2);
  int rc = orc_wsi_plane_compress(pl, tw, th, out, out_len);
  free(pl);
  return rc;
}
static int decompress_tile_blob(const u8 *blob, size_t len, int tw, int th, int channels, int bps, int ct, u8 *out) {
  if (channels == 3 && bps == 8) return orc_rgb_decompress(blob, len, tw, th, ct, out);
  size_t n = (size_t)tw * th;
  u16 *pl = (u16 *)malloc((n + 1) * sizeof(u16));
  int rc = orc_wsi_plane_decompress(blob, len, tw, th, pl);
  if (!rc) { /* uint16ToBytes :592-603 */
    if (bps <= 8) for (size_t i = 0; i < n; i++) out[i] = (u8)pl[i];
    else for (size_t i = 0; i < n; i++) { out[i * 2] = (u8)pl[i]; out[i * 2 + 1] = (u8)(pl[i] >> 8); }
  }
  free(pl);
  return rc;
}

int orc_wsi_compress(const u8 *pixels, int width, int height, int channels, int bps, int tile_w, int tile_h, int pyramid_levels,
                     u8 **out, size_t *out_len) {
  *out = NULL; *out_len = 0;
  if (tile_w == 0) tile_w = 256; /* WSIOptions.defaults, wsiformat.go:86-97 */
  if (tile_h == 0) tile_h = 256;
  int ct = 1; /* defaults() forces ColorTransform on for RGB; for grey the flag is carried but unused */
  if (channels != 3) ct = 0;
  int nlv = pyramid_levels;
  if (nlv <= 0) { /* autoLevelCount wsiformat.go:273-285 */
    nlv = 1;
    int w = width, h = height;
    while (w > tile_w || h > tile_h) { w /= 2; h /= 2; nlv++; if (w <= 1 && h <= 1) break; }
  }
  wsi_level *lv = (wsi_level *)calloc((size_t)nlv, sizeof(wsi_level));
  { /* computeLevels wsiformat.go:243-269 */
    int w = width, h = height, idx = 0;
    for (int i = 0; i < nlv; i++) {
      lv[i].w = w; lv[i].h = h;
      lv[i].tx = (w + tile_w - 1) / tile_w; lv[i].ty = (h + tile_h - 1) / tile_h;
      lv[i].first = idx;
      idx += lv[i].tx * lv[i].ty;
      w /= 2; h /= 2;
      if (w == 0) w = 1;
      if (h == 0) h = 1;
    }
  }
  int bpp = bytes_per_pixel(channels, bps);
  u8 **pyr = (u8 **)calloc((size_t)nlv, sizeof(u8 *));
  pyr[0] = (u8 *)pixels;
  for (int i = 1; i < nlv; i++) { /* wsicompress.go:45-73 */
    int pw = lv[i - 1].w, ph = lv[i - 1].h, nw, nh;
    if (i > 1) { /* previous level dims come from the downsampled image */ }
    if (pw / 2 == 0 || ph / 2 == 0) { nlv = i; break; }
    pyr[i] = (u8 *)malloc((size_t)(pw / 2) * (ph / 2) * bpp + 8);
    if (channels == 3) orc_downsample2x_rgb(pyr[i - 1], pw, ph, pyr[i], &nw, &nh);
    else if (bps <= 8) { /* bytes -> u16 -> downsample -> bytes: same arithmetic on 8-bit samples */
      size_t np = (size_t)pw * ph;
      u16 *t = (u16 *)malloc((np + 1) * 2), *d = (u16 *)malloc(((size_t)(pw / 2) * (ph / 2) + 1) * 2);
      for (size_t k = 0; k < np; k++) t[k] = pyr[i - 1][k];
      orc_downsample2x_grey(t, pw, ph, d, &nw, &nh);
      for (size_t k = 0; k < (size_t)nw * nh; k++) pyr[i][k] = (u8)d[k];
      free(t); free(d);
    } else {
      size_t np = (size_t)pw * ph;
      u16 *t = (u16 *)malloc((np + 1) * 2), *d = (u16 *)malloc(((size_t)(pw / 2) * (ph / 2) + 1) * 2);
      for (size_t k = 0; k < np; k++) t[k] = rd16(pyr[i - 1] + k * 2);
      orc_downsample2x_grey(t, pw, ph, d, &nw, &nh);
      for (size_t k = 0; k < (size_t)nw * nh; k++) { pyr[i][k * 2] = (u8)d[k]; pyr[i][k * 2 + 1] = (u8)(d[k] >> 8); }
      free(t); free(d);
    }
    lv[i].w = nw; lv[i].h = nh;
    lv[i].tx = (nw + tile_w - 1) / tile_w; lv[i].ty = (nh + tile_h - 1) / tile_h;
  }
  int total = 0;
  for (int i = 0; i < nlv; i++) { lv[i].first = total; total += lv[i].tx * lv[i].ty; }
  u8 **blobs = (u8 **)calloc((size_t)total, sizeof(u8 *));
  size_t *lens = (size_t *)calloc((size_t)total, sizeof(size_t));
  u8 *tile = (u8 *)malloc((size_t)tile_w * tile_h * bpp + 8);
  int rc = 0;
  for (int l = 0; l < nlv && !rc; l++)
    for (int ty = 0; ty < lv[l].ty && !rc; ty++)
      for (int tx = 0; tx < lv[l].tx && !rc; tx++) {
        extract_tile(pyr[l], lv[l].w, lv[l].h, tile_w, tile_h, tx, ty, bpp, tile);
        int g = lv[l].first + ty * lv[l].tx + tx;
        rc = compress_tile_blob(tile, tile_w, tile_h, channels, bps, ct, &blobs[g], &lens[g]);
      }
  free(tile);
  if (!rc) { /* WriteMIC3 wsiformat.go:99-165 */
    bbuf o = {0};
    bb_append(&o, "MIC3", 4);
    bb_u32(&o, 1); bb_u32(&o, (u32)width); bb_u32(&o, (u32)height); bb_u32(&o, (u32)tile_w); bb_u32(&o, (u32)tile_h);
    bb_push(&o, (u8)channels); bb_push(&o, (u8)(channels >> 8));
    bb_push(&o, (u8)bps);
    bb_push(&o, (u8)(0x01 | (ct ? 0x02 : 0)));
    bb_push(&o, (u8)nlv); bb_push(&o, (u8)(nlv >> 8));
    bb_push(&o, 0); bb_push(&o, 0);
    bb_u64(&o, (u64)total);
    bb_u64(&o, 0);
    for (int i = 0; i < nlv; i++) { bb_u32(&o, (u32)lv[i].w); bb_u32(&o, (u32)lv[i].h); bb_u32(&o, (u32)lv[i].tx); bb_u32(&o, (u32)lv[i].ty); bb_u32(&o, (u32)lv[i].first); }
    u64 off = 0;
    for (int g = 0; g < total; g++) { bb_u64(&o, off); bb_u64(&o, (u64)lens[g]); off += lens[g]; }
    for (int g = 0; g < total; g++) bb_append(&o, blobs[g], lens[g]);
    *out = o.p; *out_len = o.n;
  }
  for (int g = 0; g < total; g++) free(blobs[g]);
  free(blobs); free(lens);
  for (int i = 1; i < nlv; i++) free(pyr[i]);
  free(pyr); free(lv);
  return rc;
}

typedef struct {
  int width, height, tile_w, tile_h, channels, bps, ct, nlv;
  u64 total_tiles;
  size_t lv_off, table_off, data_off;
} mic3_hdr;

static int mic3_parse(const u8 *in, size_t len, mic3_hdr *h) { /* ReadMIC3Header wsiformat.go:169-227 */
  if (len < 48 || memcmp(in, "MIC3", 4) != 0) return ORC_ERR_CORRUPT;
  if (rd32(in + 4) != 1) return ORC_ERR_CORRUPT;
  h->width = (int)rd32(in + 8); h->height = (int)rd32(in + 12);
  h->tile_w = (int)rd32(in + 16); h->tile_h = (int)rd32(in + 20);
  h->channels = rd16(in + 24); h->bps = in[26]; h->ct = (in[27] & 0x02) != 0;
  h->nlv = rd16(in + 28);
  h->total_tiles = rd64(in + 32);
  h->lv_off = 48;
  if (len < h->lv_off + (size_t)h->nlv * 20) return ORC_ERR_CORRUPT;
  h->table_off = h->lv_off + (size_t)h->nlv * 20;
  if (h->total_tiles > (len - h->table_off) / 16) return ORC_ERR_CORRUPT;
  h->data_off = h->table_off + (size_t)h->total_tiles * 16;
  return 0;
}
static wsi_level mic3_level(const u8 *in, const mic3_hdr *h, int l) {
  const u8 *p = in + h->lv_off + (size_t)l * 20;
  wsi_level v = {(int)rd32(p), (int)rd32(p + 4), (int)rd32(p + 8), (int)rd32(p + 12), (int)rd32(p + 16)};
  return v;
}
static int mic3_tile(const u8 *in, size_t len, const mic3_hdr *h, int idx, const u8 **p, size_t *l) {
  if (idx < 0 || (u64)idx >= h->total_tiles) return ORC_ERR_ARG;
  u64 o = rd64(in + h->table_off + (size_t)idx * 16), n = rd64(in + h->table_off + (size_t)idx * 16 + 8);
  if (h->data_off + o + n > len) return ORC_ERR_CORRUPT;
  *p = in + h->data_off + o; *l = (size_t)n;
  return 0;
}

int orc_wsi_header(const u8 *in, size_t len, int *width, int *height, int *tile_w, int *tile_h, int *channels, int *bps,
                   int *color_transform, int *nlevels, int *level_info, int max_levels, u64 *total_tiles) {
  mic3_hdr h;
  int rc = mic3_parse(in, len, &h);
  if (rc) return rc;
  *width = h.width; *height = h.height; *tile_w = h.tile_w; *tile_h = h.tile_h; *channels = h.channels; *bps = h.bps;
  *color_transform = h.ct; *nlevels = h.nlv; *total_tiles = h.total_tiles;
  for (int l = 0; l < h.nlv && l < max_levels; l++) {
    wsi_level v = mic3_level(in, &h, l);
    level_info[l * 5] = v.w; level_info[l * 5 + 1] = v.h; level_info[l * 5 + 2] = v.tx; level_info[l * 5 + 3] = v.ty; level_info[l * 5 + 4] = v.first;
  }
  return 0;
}

int orc_wsi_decompress_tile(const u8 *in, size_t len, int level, int tx, int ty, u8 **out, size_t *out_len, int *tw, int *th) { /* :175-217 */
  *out = NULL; *out_len = 0;
  mic3_hdr h;
  int rc = mic3_parse(in, len, &h);
  if (rc) return rc;
  if (level < 0 || level >= h.nlv) return ORC_ERR_ARG;
  wsi_level lv = mic3_level(in, &h, level);
  if (tx < 0 || tx >= lv.tx || ty < 0 || ty >= lv.ty) return ORC_ERR_ARG;
  const u8 *blob; size_t bl;
  rc = mic3_tile(in, len, &h, lv.first + ty * lv.tx + tx, &blob, &bl);
  if (rc) return rc;
  int bpp = bytes_per_pixel(h.channels, h.bps);
  u8 *tile = (u8 *)malloc((size_t)h.tile_w * h.tile_h * bpp + 8);
  rc = decompress_tile_blob(blob, bl, h.tile_w, h.tile_h, h.channels, h.bps, h.ct, tile);
  if (rc) { free(tile); return rc; }
  int aw = h.tile_w, ah = h.tile_h;
  int ex = lv.w - tx * h.tile_w, ey = lv.h - ty * h.tile_h;
  if (ex < aw) aw = ex;
  if (ey < ah) ah = ey;
  if (aw != h.tile_w || ah != h.tile_h) { /* cropTile :558-572 */
    u8 *c = (u8 *)malloc((size_t)aw * ah * bpp + 8);
    for (int y = 0; y < ah; y++) memcpy(c + (size_t)y * aw * bpp, tile + (size_t)y * h.tile_w * bpp, (size_t)aw * bpp);
    free(tile);
    tile = c;
  }
  *out = tile; *out_len = (size_t)aw * ah * bpp; *tw = aw; *th = ah;
  return 0;
}

int orc_wsi_decompress_region(const u8 *in, size_t len, int level, int x, int y, int w, int hgt, u8 **out, size_t *out_len, int *ow, int *oh) { /* :220-296 */
  *out = NULL; *out_len = 0;
  mic3_hdr h;
  int rc = mic3_parse(in, len, &h);
  if (rc) return rc;
  if (level < 0 || level >= h.nlv) return ORC_ERR_ARG;
  wsi_level lv = mic3_level(in, &h, level);
  if (x + w > lv.w) w = lv.w - x;
  if (y + hgt > lv.h) hgt = lv.h - y;
  if (w <= 0 || hgt <= 0) return ORC_ERR_ARG;
  int bpp = bytes_per_pixel(h.channels, h.bps);
  int stx = x / h.tile_w, sty = y / h.tile_h, etx = (x + w - 1) / h.tile_w, ety = (y + hgt - 1) / h.tile_h;
  u8 *res = (u8 *)calloc((size_t)w * hgt * bpp + 8, 1);
  u8 *tile = (u8 *)malloc((size_t)h.tile_w * h.tile_h * bpp + 8);
  for (int ty = sty; ty <= ety && !rc; ty++)
    for (int tx = stx; tx <= etx && !rc; tx++) {
      const u8 *blob; size_t bl;
      rc = mic3_tile(in, len, &h, lv.first + ty * lv.tx + tx, &blob, &bl);
      if (rc) break;
      rc = decompress_tile_blob(blob, bl, h.tile_w, h.tile_h, h.channels, h.bps, h.ct, tile);
      if (rc) break;
      int tsx = tx * h.tile_w, tsy = ty * h.tile_h, tw = h.tile_w, th = h.tile_h;
      if (tsx + tw > lv.w) tw = lv.w - tsx;
      if (tsy + th > lv.h) th = lv.h - tsy;
      int ox0 = x > tsx ? x : tsx, oy0 = y > tsy ? y : tsy;
      int ox1 = (x + w) < (tsx + tw) ? (x + w) : (tsx + tw), oy1 = (y + hgt) < (tsy + th) ? (y + hgt) : (tsy + th);
      for (int ry = oy0; ry < oy1; ry++) {
        size_t so = ((size_t)(ry - tsy) * h.tile_w + (ox0 - tsx)) * bpp;
        size_t dof = ((size_t)(ry - y) * w + (ox0 - x)) * bpp;
        memcpy(res + dof, tile + so, (size_t)(ox1 - ox0) * bpp);
      }
    }
  free(tile);
  if (rc) { free(res); return rc; }
  *out = res; *out_len = (size_t)w * hgt * bpp; *ow = w; *oh = hgt;
  return 0;
}
