/* synth.c -- deterministic synthetic inputs for tests and bench.
 *
 * PRNG: xorshift64* seeded 0x5A50415142323030 + block_index, so that any host (C, C#, Python via
 * this library) regenerates the same bytes for block i without generating the blocks before it.
 *
 *   text : order-2 word Markov chain over a fixed 4096-word vocabulary (Zipf-ranked choices,
 *          words of 2..12 lower-case letters, spaces / punctuation / newlines, lines ~72 wide)
 *   mixed: per 64 KB chunk one of: text (50 %), little-endian records with a counter, a slowly
 *          increasing stamp, 8 random bytes and 16 bytes repeated from 16 records earlier (20 %),
 *          uniform random bytes (15 %), x86-like stream with E8 xx xx xx 00|FF calls (15 %)
 */
#include <stdint.h>
#include <string.h>
#include <stdlib.h>

static inline uint64_t xs(uint64_t* s) {
  uint64_t x = *s;
  x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
  *s = x;
  return x * 0x2545F4914F6CDD1DULL;
}

#define NVOCAB 4096
static char g_word[NVOCAB][13];
static uint8_t g_len[NVOCAB];
static int g_ready = 0;

static void build_vocab(void) {
  if (g_ready) return;
  uint64_t s = 0x5A50415142323030ULL ^ 0x766F636162ULL;
  static const char* freq = "eeeeeeeeeeeetttttttttaaaaaaaaooooooooiiiiiiinnnnnnnsssssshhhhhhrrrrrrddddllllcccuuummmwwffggyyppbbvkjxqz";
  const int nf = (int)strlen(freq);
  for (int i = 0; i < NVOCAB; ++i) {
    int len = 2 + (int)(xs(&s) % 11);
    if (i < 64) len = 1 + (int)(xs(&s) % 4) + 1;
    for (int k = 0; k < len; ++k) g_word[i][k] = freq[xs(&s) % nf];
    g_word[i][len] = 0;
    g_len[i] = (uint8_t)len;
  }
  g_ready = 1;
}

/* Zipf-like rank in [0, n): most draws land on small ranks. */
static inline uint32_t zipf(uint64_t* s, uint32_t n) {
  uint64_t r = xs(s);
  uint32_t bits = (uint32_t)(r & 63);
  uint32_t span = 1;
  while (span < n && (bits & 1)) { span <<= 1; bits >>= 1; }
  span <<= 2;
  if (span > n) span = n;
  return (uint32_t)((r >> 8) % span);
}

static size_t gen_text(uint64_t* s, uint8_t* out, size_t n, uint32_t* w1, uint32_t* w2, int* col) {
  size_t pos = 0;
  while (pos < n) {
    /* the successor list of (w1,w2) is a pseudo-random permutation window of the vocabulary */
    uint32_t ctx = (*w1 * 2654435761u) ^ (*w2 * 40503u);
    uint32_t rank = zipf(s, 256);
    uint32_t w = (ctx + rank * 0x9E3779B1u) & (NVOCAB - 1);
    if ((xs(s) & 7) == 0) w = zipf(s, NVOCAB);
    const char* word = g_word[w];
    int len = g_len[w];
    for (int k = 0; k < len && pos < n; ++k) out[pos++] = (uint8_t)word[k];
    *col += len;
    uint64_t r = xs(s);
    if (pos < n) {
      unsigned sel = (unsigned)(r & 63);
      if (sel == 0 && pos + 1 < n) { out[pos++] = '.'; out[pos++] = ' '; *col += 2; }
      else if (sel == 1 && pos + 1 < n) { out[pos++] = ','; out[pos++] = ' '; *col += 2; }
      else if (*col > 64 + (int)((r >> 8) & 15)) { out[pos++] = '\n'; *col = 0; }
      else { out[pos++] = ' '; *col += 1; }
    }
    *w1 = *w2; *w2 = w;
  }
  return pos;
}

void synth_text_block(uint64_t block_index, uint8_t* out, uint64_t n) {
  build_vocab();
  uint64_t s = 0x5A50415142323030ULL + block_index;
  uint32_t w1 = 0, w2 = 0; int col = 0;
  gen_text(&s, out, (size_t)n, &w1, &w2, &col);
}

void synth_mixed_block(uint64_t block_index, uint8_t* out, uint64_t n) {
  build_vocab();
  uint64_t s = 0x5A50415142323030ULL + block_index;
  uint32_t w1 = 0, w2 = 0; int col = 0;
  uint32_t counter = (uint32_t)(block_index * 1000003u), stamp = 0x5F000000u + (uint32_t)block_index * 977u;
  uint64_t pos = 0;
  while (pos < n) {
    uint64_t len = n - pos < 65536 ? n - pos : 65536;
    uint8_t* p = out + pos;
    unsigned kind = (unsigned)(xs(&s) % 100);
    if (kind < 50) gen_text(&s, p, (size_t)len, &w1, &w2, &col);
    else if (kind < 70) {
      uint64_t q = 0;
      while (q < len) {
        uint8_t rec[32];
        memcpy(rec, &counter, 4); ++counter;
        stamp += (uint32_t)(xs(&s) & 1023); memcpy(rec + 4, &stamp, 4);
        uint64_t r = xs(&s); memcpy(rec + 8, &r, 8);
        if (q >= 16 * 32) memcpy(rec + 16, p + q - 16 * 32 + 16, 16);
        else { uint64_t a = xs(&s), b = xs(&s); memcpy(rec + 16, &a, 8); memcpy(rec + 24, &b, 8); }
        uint64_t k = len - q < 32 ? len - q : 32;
        memcpy(p + q, rec, k);
        q += k;
      }
    } else if (kind < 85) {
      uint64_t q = 0;
      while (q + 8 <= len) { uint64_t r = xs(&s); memcpy(p + q, &r, 8); q += 8; }
      while (q < len) p[q++] = (uint8_t)xs(&s);
    } else {
      uint64_t q = 0;
      static const uint8_t ops[16] = {0x8b, 0x89, 0x55, 0x48, 0x83, 0xc3, 0x90, 0x0f, 0x85, 0x74, 0x75, 0xff, 0x50, 0x5d, 0x31, 0xc0};
      while (q < len) {
        unsigned gap = 5 + (unsigned)(xs(&s) % 36);
        for (unsigned k = 0; k < gap && q < len; ++k) {
          uint64_t r = xs(&s);
          p[q++] = (r & 3) ? ops[(r >> 2) & 15] : (uint8_t)(r >> 8);
        }
        if (q + 5 <= len) {
          uint64_t r = xs(&s);
          uint32_t target = (uint32_t)((r >> 8) % 0x40000);           /* a few hundred hot call targets */
          target = (target & ~0xFFFu) | ((target * 2654435761u >> 20) & 0xFF0u);
          uint32_t rel = target - (uint32_t)(pos + q);                /* relative displacement */
          p[q++] = (r & 7) ? 0xE8 : 0xE9;
          p[q++] = (uint8_t)rel; p[q++] = (uint8_t)(rel >> 8); p[q++] = (uint8_t)(rel >> 16);
          p[q++] = (rel & 0x800000u) ? 0xFF : 0x00;
        }
      }
    }
    pos += len;
  }
}

/* Fill `count` consecutive blocks of `block_bytes` each, starting at block `first`. */
void synth_fill(int mixed, uint64_t first, uint64_t count, uint64_t block_bytes, uint8_t* out) {
  for (uint64_t i = 0; i < count; ++i) {
    if (mixed) synth_mixed_block(first + i, out + i * block_bytes, block_bytes);
    else synth_text_block(first + i, out + i * block_bytes, block_bytes);
  }
}
