/* libb200st_io: TFRecord framing + CRC-32C on the host (see include/b200st_io.h).  Plain C, built with gcc. */
#include "../../include/b200st_io.h"

#include <string.h>

int b200st_io_version(void) { return 2; }

/* slicing-by-8 tables for the reflected Castagnoli polynomial, built on first use */
static uint32_t g_tab[8][256];
static int g_tab_ready = 0;

uint32_t b200st_crc32c_table(uint32_t crc, const void* data, size_t n);     /* portable path (also the cross-check in the tests) */

static void build_tables(void) {
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ 0x82F63B78u : (c >> 1);
    g_tab[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 8; ++t) g_tab[t][i] = (g_tab[t - 1][i] >> 8) ^ g_tab[0][g_tab[t - 1][i] & 0xffu];
  __atomic_store_n(&g_tab_ready, 1, __ATOMIC_RELEASE);
}

#if defined(__x86_64__)
/* SSE4.2 crc32 instruction (the Castagnoli polynomial is the one it implements): ~3x the table version on one stream */
__attribute__((target("sse4.2"))) static uint32_t crc32c_hw(uint32_t crc, const uint8_t* p, size_t n) {
  uint64_t c = (uint32_t)~crc;
  while (n && ((uintptr_t)p & 7u)) { c = __builtin_ia32_crc32qi((uint32_t)c, *p++); --n; }
  while (n >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    c = __builtin_ia32_crc32di(c, w);
    p += 8; n -= 8;
  }
  while (n--) c = __builtin_ia32_crc32qi((uint32_t)c, *p++);
  return ~(uint32_t)c;
}
#endif

uint32_t b200st_crc32c(uint32_t crc, const void* data, size_t n) {
#if defined(__x86_64__)
  static int hw = -1;
  if (hw < 0) hw = __builtin_cpu_supports("sse4.2") ? 1 : 0;
  if (hw) return crc32c_hw(crc, (const uint8_t*)data, n);
#endif
  return b200st_crc32c_table(crc, data, n);
}

uint32_t b200st_crc32c_table(uint32_t crc, const void* data, size_t n) {
  if (!__atomic_load_n(&g_tab_ready, __ATOMIC_ACQUIRE)) build_tables();      /* idempotent: a race only rebuilds equal tables */
  const uint8_t* p = (const uint8_t*)data;
  uint32_t c = ~crc;
  while (n && ((uintptr_t)p & 7u)) { c = (c >> 8) ^ g_tab[0][(c ^ *p++) & 0xffu]; --n; }
  while (n >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    w ^= c;                                                                  /* little-endian host (x86-64 / aarch64) */
    c = g_tab[7][w & 0xff] ^ g_tab[6][(w >> 8) & 0xff] ^ g_tab[5][(w >> 16) & 0xff] ^ g_tab[4][(w >> 24) & 0xff] ^
        g_tab[3][(w >> 32) & 0xff] ^ g_tab[2][(w >> 40) & 0xff] ^ g_tab[1][(w >> 48) & 0xff] ^ g_tab[0][(w >> 56) & 0xff];
    p += 8; n -= 8;
  }
  while (n--) c = (c >> 8) ^ g_tab[0][(c ^ *p++) & 0xffu];
  return ~c;
}

uint32_t b200st_crc32c_mask(uint32_t crc) { return ((crc >> 15) | (crc << 17)) + 0xa282ead8u; }

static uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }

int64_t b200st_tfrecord_index(const void* buf, size_t n, int64_t* offsets, int64_t* lengths, int64_t max_records, int verify) {
  const uint8_t* b = (const uint8_t*)buf;
  size_t pos = 0;
  int64_t count = 0;
  while (pos < n) {
    if (n - pos < 12) return -1 - (int64_t)pos;
    uint64_t len = rd64(b + pos);
    if (verify >= 1 && rd32(b + pos + 8) != b200st_crc32c_mask(b200st_crc32c(0, b + pos, 8))) return -1 - (int64_t)pos;
    if (len > n - pos - 12 || n - pos - 12 - len < 4) return -1 - (int64_t)pos;
    if (verify >= 2 && rd32(b + pos + 12 + len) != b200st_crc32c_mask(b200st_crc32c(0, b + pos + 12, (size_t)len)))
      return -1 - (int64_t)pos;
    if (count < max_records) {
      if (offsets) offsets[count] = (int64_t)(pos + 12);
      if (lengths) lengths[count] = (int64_t)len;
    }
    ++count;
    pos += 12 + (size_t)len + 4;
  }
  return count;
}

void b200st_tfrecord_frame(const void* payload, size_t n, uint8_t* header12, uint8_t* footer4) {
  uint64_t len = (uint64_t)n;
  memcpy(header12, &len, 8);
  uint32_t c = b200st_crc32c_mask(b200st_crc32c(0, header12, 8));
  memcpy(header12 + 8, &c, 4);
  c = b200st_crc32c_mask(b200st_crc32c(0, payload, n));
  memcpy(footer4, &c, 4);
}

/* ---------------------------------------------------------------------------------------------- tf.train.Example */
static int rd_varint(const uint8_t* p, size_t n, size_t* pos, uint64_t* out) {
  uint64_t x = 0;
  int shift = 0;
  while (*pos < n) {
    const uint8_t b = p[(*pos)++];
    if (shift < 64) x |= (uint64_t)(b & 0x7f) << shift;
    if (b < 0x80) { *out = x; return 0; }
    shift += 7;
    if (shift > 63) return -1;
  }
  return -1;
}
/* next field of the message p[*pos, end): number, wire type, and for length-delimited fields the payload range */
static int next_field(const uint8_t* p, size_t end, size_t* pos, uint32_t* num, uint32_t* wt, size_t* a, size_t* b, uint64_t* val) {
  uint64_t key;
  if (rd_varint(p, end, pos, &key)) return -1;
  *num = (uint32_t)(key >> 3); *wt = (uint32_t)(key & 7);
  if (*wt == 0) return rd_varint(p, end, pos, val);
  if (*wt == 2) {
    uint64_t l;
    if (rd_varint(p, end, pos, &l) || l > end - *pos) return -1;
    *a = *pos; *b = *pos + (size_t)l; *pos = *b;
    return 0;
  }
  if (*wt == 5) { if (end - *pos < 4) return -1; *a = *pos; *b = *pos + 4; *pos += 4; return 0; }
  if (*wt == 1) { if (end - *pos < 8) return -1; *a = *pos; *b = *pos + 8; *pos += 8; return 0; }
  return -1;
}

static int64_t count_varints(const uint8_t* p, size_t a, size_t b) {
  int64_t c = 0;
  for (size_t i = a; i < b; ++i) c += p[i] < 0x80;
  return (b > a && p[b - 1] >= 0x80) ? -1 : c;
}

int b200st_example_lookup(const void* rec, size_t n, const char* const* keys, int nkeys, int32_t* kind, int64_t* off, int64_t* len,
                          int64_t* count) {
  const uint8_t* p = (const uint8_t*)rec;
  for (int i = 0; i < nkeys; ++i) { kind[i] = B200ST_FEATURE_NONE; off[i] = 0; len[i] = 0; count[i] = 0; }
  size_t pos = 0, a, b;
  uint32_t num, wt;
  uint64_t v;
  while (pos < n) {                                   /* Example */
    if (next_field(p, n, &pos, &num, &wt, &a, &b, &v)) return -1;
    if (num != 1 || wt != 2) continue;
    size_t fpos = a;
    const size_t fend = b;
    while (fpos < fend) {                             /* Features: map entries */
      size_t ea, eb;
      if (next_field(p, fend, &fpos, &num, &wt, &ea, &eb, &v)) return -1;
      if (num != 1 || wt != 2) continue;
      size_t epos = ea, ka = 0, kb = 0, va = 0, vb = 0;
      int has_v = 0;
      while (epos < eb) {                             /* entry: key = 1, value = 2 */
        size_t xa, xb;
        if (next_field(p, eb, &epos, &num, &wt, &xa, &xb, &v)) return -1;
        if (num == 1 && wt == 2) { ka = xa; kb = xb; }
        else if (num == 2 && wt == 2) { va = xa; vb = xb; has_v = 1; }
      }
      int which = -1;
      for (int i = 0; i < nkeys; ++i)
        if (strlen(keys[i]) == kb - ka && memcmp(keys[i], p + ka, kb - ka) == 0) { which = i; break; }
      if (which < 0 || !has_v) continue;
      size_t vpos = va;
      while (vpos < vb) {                             /* Feature: oneof 1 bytes_list, 2 float_list, 3 int64_list */
        size_t la, lb;
        if (next_field(p, vb, &vpos, &num, &wt, &la, &lb, &v)) return -1;
        if (wt != 2 || num < 1 || num > 3) continue;
        kind[which] = (int32_t)num;
        off[which] = 0; len[which] = 0; count[which] = 0;
        size_t lpos = la;
        int chunks = 0;
        while (lpos < lb) {                           /* the list: repeated field 1 (packed = one length-delimited chunk) */
          size_t ca, cb;
          uint32_t n2, w2;
          if (next_field(p, lb, &lpos, &n2, &w2, &ca, &cb, &v)) return -1;
          if (n2 != 1) continue;
          if (num == 1) {                             /* bytes values: remember the first, count them all */
            if (w2 != 2) return -1;
            if (count[which] == 0) { off[which] = (int64_t)ca; len[which] = (int64_t)(cb - ca); }
            count[which] += 1;
          } else if (w2 == 2 && chunks == 0) {
            off[which] = (int64_t)ca; len[which] = (int64_t)(cb - ca);
            if (num == 2) { if ((cb - ca) % 4) return -1; count[which] = (int64_t)((cb - ca) / 4); }
            else { count[which] = count_varints(p, ca, cb); if (count[which] < 0) return -1; }
            chunks = 1;
          } else {
            count[which] = -1;                        /* unpacked scalars or several chunks: general decoder */
            chunks = 2;
          }
        }
        if (count[which] == 0) kind[which] = B200ST_FEATURE_NONE;      /* an empty list reads like a missing feature */
      }
    }
  }
  return pos == n ? 0 : -1;
}

int64_t b200st_decode_varints(const void* src, size_t n, int64_t* out, int64_t count) {
  const uint8_t* p = (const uint8_t*)src;
  size_t pos = 0;
  int64_t i = 0;
  while (i < count && pos < n) {
    uint64_t v;
    if (rd_varint(p, n, &pos, &v)) break;
    out[i++] = (int64_t)v;
  }
  return i;
}

void b200st_pad_rows_f32(float* dst, int64_t row_elems, const float* const* src, const int64_t* n_elems, int32_t B) {
  for (int32_t b = 0; b < B; ++b) {
    int64_t n = n_elems[b] < row_elems ? n_elems[b] : row_elems;
    if (n < 0) n = 0;
    float* row = dst + (int64_t)b * row_elems;
    if (n) memcpy(row, src[b], (size_t)n * sizeof(float));
    if (n < row_elems) memset(row + n, 0, (size_t)(row_elems - n) * sizeof(float));
  }
}
