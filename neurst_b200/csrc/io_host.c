/* libb200st_io: TFRecord framing + CRC-32C on the host (see include/b200st_io.h).  Plain C, built with gcc. */
#include "../../include/b200st_io.h"

#include <string.h>

int b200st_io_version(void) { return 1; }

/* slicing-by-8 tables for the reflected Castagnoli polynomial, built on first use */
static uint32_t g_tab[8][256];
static int g_tab_ready = 0;

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

uint32_t b200st_crc32c(uint32_t crc, const void* data, size_t n) {
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
