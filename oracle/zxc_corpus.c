/*
 * zxc_corpus.c -- TEST / BENCH INFRASTRUCTURE ONLY (not part of libzxc).
 *
 * Deterministic "Silesia-shaped" synthetic corpus.
 * There is no Silesia in the container and no network, so the benchmark input
 * is synthesised by data class with Silesia's size proportions:
 *
 *   dickens 10.2  mozilla 51.2  mr 10.0  nci 33.6  ooffice 6.2  osdb 10.1
 *   reymont 6.6   samba 21.6    sao 7.3  webster 41.5  xml 5.3  x-ray 8.5   (MB, 211.9 total)
 *
 * The stream is defined chunk-wise: chunk c (1 MiB) is a pure function of
 * (seed, c), so any byte range can be produced by any number of threads and
 * tiles of the 212-chunk unit never repeat (the chunk index salts the RNG).
 */
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define CHUNK ((size_t)1 << 20)
#define UNIT_CHUNKS 212

typedef struct { uint64_t s; } rng_t;
static inline uint64_t rnd(rng_t* r) { /* xorshift64* */
    uint64_t x = r->s;
    x ^= x >> 12;
    x ^= x << 25;
    x ^= x >> 27;
    r->s = x;
    return x * 0x2545F4914F6CDD1Dull;
}
static inline uint32_t rndn(rng_t* r, uint32_t n) { return (uint32_t)((rnd(r) >> 32) * (uint64_t)n >> 32); }
/* Zipf-ish index in [0, n): squares a uniform variate, biasing towards 0 */
static inline uint32_t zipf(rng_t* r, uint32_t n) {
    const uint64_t u = rnd(r) >> 40; /* 24 bits */
    const uint64_t v = (u * u) >> 24;
    return (uint32_t)((v * v >> 24) * n >> 24);
}

enum { K_TEXT_EN, K_TEXT_PL, K_TEXT_DICT, K_EXE, K_IMG, K_CHEM, K_DB, K_SRC, K_FLOAT, K_XML };

/* class of chunk u within the 212-chunk unit, in Silesia file order */
static int unit_class(unsigned u) {
    static const struct { unsigned mb; int k; } seg[12] = {
        {10, K_TEXT_EN}, {51, K_EXE}, {10, K_IMG}, {34, K_CHEM}, {6, K_EXE}, {10, K_DB},
        {7, K_TEXT_PL},  {22, K_SRC}, {7, K_FLOAT}, {41, K_TEXT_DICT}, {5, K_XML}, {9, K_IMG}};
    unsigned acc = 0;
    for (int i = 0; i < 12; i++) {
        acc += seg[i].mb;
        if (u < acc) return seg[i].k;
    }
    return K_IMG;
}

/* ---- vocabulary shared by the text-like classes ---- */
#define VOCAB 4096
typedef struct { char w[VOCAB][12]; uint8_t len[VOCAB]; uint16_t next[VOCAB][4]; } vocab_t;

static void vocab_build(vocab_t* v, uint64_t seed, const char* alphabet, unsigned alen) {
    rng_t r = {seed | 1};
    for (int i = 0; i < VOCAB; i++) {
        unsigned L = 2 + rndn(&r, 3) + (i > 64 ? rndn(&r, 4) : 0) + (i > 1024 ? rndn(&r, 3) : 0);
        if (L > 11) L = 11;
        for (unsigned k = 0; k < L; k++) {
            /* letter frequencies skewed towards the front of the alphabet */
            unsigned a = rndn(&r, alen), b = rndn(&r, alen);
            v->w[i][k] = alphabet[a < b ? a : b];
        }
        v->len[i] = (uint8_t)L;
        for (int k = 0; k < 4; k++) v->next[i][k] = (uint16_t)zipf(&r, VOCAB);
    }
}

static size_t put(uint8_t* d, size_t p, size_t n, const void* s, size_t len) {
    if (p + len > n) len = n - p;
    memcpy(d + p, s, len);
    return p + len;
}

static void gen_text(uint8_t* d, size_t n, rng_t* r, const vocab_t* v, int dict_like) {
    size_t p = 0;
    uint32_t w = zipf(r, VOCAB);
    int sent = 0;
    while (p < n) {
        const uint32_t x = (uint32_t)(rnd(r) >> 32);
        if ((x & 15) < 12) w = v->next[w][((x >> 4) & 3) & ((x >> 6) & 3)]; else w = zipf(r, VOCAB);
        char tmp[16];
        unsigned L = v->len[w];
        memcpy(tmp, v->w[w], L);
        if (sent == 0 && tmp[0] >= 'a' && tmp[0] <= 'z') tmp[0] -= 32;
        p = put(d, p, n, tmp, L);
        sent++;
        const uint32_t y = x >> 8;
        if (sent > 6 && (y & 7) == 0) {
            p = put(d, p, n, (y & 64) ? "? " : ". ", 2);
            sent = 0;
            if ((y & 0x700) == 0) p = put(d, p, n, "\n\n", dict_like ? 1 : 2);
            if (dict_like && (y & 0x300) == 0) { /* headword line, as in a dictionary */
                const uint32_t h = zipf(r, VOCAB);
                char up[16];
                for (unsigned k = 0; k < v->len[h]; k++) up[k] = (char)(v->w[h][k] & ~32);
                p = put(d, p, n, "\n", 1);
                p = put(d, p, n, up, v->len[h]);
                p = put(d, p, n, ", n. [", 6);
            }
        } else if ((y & 31) == 1) {
            p = put(d, p, n, ", ", 2);
        } else {
            p = put(d, p, n, " ", 1);
        }
    }
}

static void gen_exe(uint8_t* d, size_t n, rng_t* r, uint64_t salt) {
    /* 256 "instruction" patterns of 1..7 bytes with skewed use, small immediates, aligned
     * pointers sharing high bytes, zero / 0xCC padding runs, and short repeated functions */
    uint8_t pat[256][8];
    uint8_t plen[256];
    rng_t pr = {0x9E3779B97F4A7C15ull ^ (salt % 7)};
    for (int i = 0; i < 256; i++) {
        plen[i] = (uint8_t)(1 + rndn(&pr, 6));
        for (int k = 0; k < 8; k++) pat[i][k] = (uint8_t)(rnd(&pr) >> 56);
    }
    size_t p = 0;
    const uint32_t base = 0x00400000u + (rndn(r, 64) << 16);
    while (p < n) {
        const uint32_t x = (uint32_t)(rnd(r) >> 32);
        const unsigned sel = x & 63;
        if (sel < 40) {
            const unsigned i = zipf(r, 256);
            p = put(d, p, n, pat[i], plen[i]);
            if ((x >> 8 & 3) == 0) { /* imm8 */
                uint8_t b = (uint8_t)(zipf(r, 64) * 4);
                p = put(d, p, n, &b, 1);
            }
        } else if (sel < 50) { /* call/jmp rel32 or absolute pointer */
            uint8_t op = (x >> 8 & 1) ? 0xE8 : 0x8B;
            uint32_t a = (x >> 9 & 1) ? base + (zipf(r, 4096) << 4) : (uint32_t)(-(int32_t)zipf(r, 65536));
            p = put(d, p, n, &op, 1);
            p = put(d, p, n, &a, 4);
        } else if (sel < 54) { /* zero / int3 padding to 16 */
            uint8_t fill[16];
            memset(fill, (x >> 8 & 1) ? 0xCC : 0, 16);
            p = put(d, p, n, fill, 16 - (p & 15));
        } else if (sel < 60 && p > 4096) { /* a copy of an earlier function body */
            const size_t len = 24 + rndn(r, 400), back = 64 + rndn(r, p > 60000 ? 60000 : (uint32_t)p - 64);
            for (size_t k = 0; k < len && p < n; k++, p++) d[p] = d[p - back];
        } else if (sel < 62) { /* data table: 4-byte entries with a constant stride */
            uint32_t v0 = base + (rndn(r, 1 << 16) << 2);
            const unsigned cnt = 4 + rndn(r, 40), st = 4u << rndn(r, 4);
            for (unsigned k = 0; k < cnt; k++, v0 += st) p = put(d, p, n, &v0, 4);
        } else { /* string table entry */
            static const char* s[8] = {"GetProcAddress", "%s: error %d\n", "kernel32.dll", "assertion failed",
                                       "\0\0\0\0", "Microsoft", "__cxa_", "operator new"};
            const char* t = s[x >> 8 & 7];
            p = put(d, p, n, t, strlen(t) + 1);
        }
    }
}

static void gen_img(uint8_t* d, size_t n, rng_t* r) {
    /* 12-bit samples, little-endian u16: smooth field + noise; rows of 1024 samples */
    int32_t v = 1800 + (int32_t)rndn(r, 400), slope = 0;
    for (size_t p = 0; p + 1 < n; p += 2) {
        const uint32_t x = (uint32_t)(rnd(r) >> 32);
        if ((x & 63) == 0) slope = (int32_t)(x >> 8 & 15) - 7;
        v += slope + (int32_t)((x >> 12 & 31) + (x >> 17 & 31)) - 31;
        if (v < 0) v = 0;
        if (v > 4095) v = 4095;
        if ((x >> 24) < 40) { /* flat background runs */
            size_t run = 8 + (x >> 22 & 63);
            for (; run && p + 1 < n; run--, p += 2) { d[p] = (uint8_t)v; d[p + 1] = (uint8_t)(v >> 8); }
            if (p + 1 >= n) break;
        }
        d[p] = (uint8_t)v;
        d[p + 1] = (uint8_t)(v >> 8);
    }
    if (n & 1) d[n - 1] = 0;
}

static size_t put_num(uint8_t* d, size_t p, size_t n, uint32_t v, int width) {
    char t[16];
    int L = snprintf(t, sizeof t, "%*u", width, v);
    return put(d, p, n, t, (size_t)L);
}

static void gen_chem(uint8_t* d, size_t n, rng_t* r) {
    /* MDL-molfile-like records: coordinate lines and bond lines, very repetitive */
    static const char* el[8] = {"C  ", "C  ", "C  ", "H  ", "O  ", "N  ", "C  ", "S  "};
    size_t p = 0;
    uint32_t id = rndn(r, 100000);
    while (p < n) {
        const unsigned atoms = 8 + rndn(r, 30);
        p = put_num(d, p, n, id++, 7);
        p = put(d, p, n, "\n  -ISIS-  \n\n", 13);
        p = put_num(d, p, n, atoms, 3);
        p = put_num(d, p, n, atoms + rndn(r, 3), 3);
        p = put(d, p, n, "  0  0  0  0  0  0  0  0999 V2000\n", 35);
        for (unsigned a = 0; a < atoms && p < n; a++) {
            for (int c = 0; c < 3; c++) {
                char t[16];
                const int L = snprintf(t, sizeof t, "%5d.%04u", (int)rndn(r, 12) - 6, c == 2 ? 0u : rndn(r, 100) * 100);
                p = put(d, p, n, t, (size_t)L);
            }
            p = put(d, p, n, " ", 1);
            p = put(d, p, n, el[rndn(r, 8)], 3);
            p = put(d, p, n, " 0  0  0  0  0  0  0  0  0  0  0  0\n", 37);
        }
        for (unsigned b = 1; b < atoms && p < n; b++) {
            p = put_num(d, p, n, b, 3);
            p = put_num(d, p, n, b + 1 - (rndn(r, 4) == 0 ? rndn(r, b) : 0), 3);
            p = put_num(d, p, n, 1 + (rndn(r, 5) == 0), 3);
            p = put(d, p, n, "  0  0  0  0\n", 13);
        }
        p = put(d, p, n, "M  END\n$$$$\n", 12);
    }
}

static void gen_db(uint8_t* d, size_t n, rng_t* r, const vocab_t* v) {
    /* fixed-width rows: ascending key, numeric fields, padded strings from a small vocabulary */
    size_t p = 0;
    uint32_t key = rndn(r, 1 << 20);
    while (p < n) {
        uint8_t row[96];
        memset(row, ' ', sizeof row);
        key += 1 + rndn(r, 3);
        memcpy(row, &key, 4);
        uint32_t f1 = rndn(r, 1000), f2 = zipf(r, 50000);
        memcpy(row + 4, &f1, 4);
        memcpy(row + 8, &f2, 4);
        uint64_t ts = 0x5F000000ull + key * 37ull;
        memcpy(row + 12, &ts, 8);
        for (int s = 0, o = 20; s < 3; s++, o += 24) {
            const uint32_t w = zipf(r, s == 0 ? 256 : 2048);
            memcpy(row + o, v->w[w], v->len[w]);
            if (s == 2) {
                const uint32_t w2 = zipf(r, 512);
                memcpy(row + o + v->len[w] + 1, v->w[w2], v->len[w2]);
            }
        }
        row[92] = (uint8_t)rndn(r, 4);
        row[93] = row[94] = row[95] = 0;
        p = put(d, p, n, row, sizeof row);
    }
}

static void gen_src(uint8_t* d, size_t n, rng_t* r, const vocab_t* v) {
    static const char* kw[16] = {"if (", "return ", "static int ", "struct ", "for (i = 0; i < ", "#include <",
                                 "    ", "\t", "} else {", "NULL", "const char *", "->", " = ", "();\n", "/* ", " */\n"};
    size_t p = 0;
    while (p < n) {
        const uint32_t x = (uint32_t)(rnd(r) >> 32);
        const unsigned indent = x & 3;
        for (unsigned k = 0; k < indent; k++) p = put(d, p, n, "\t", 1);
        const unsigned items = 2 + (x >> 2 & 7);
        for (unsigned k = 0; k < items && p < n; k++) {
            const uint32_t y = (uint32_t)(rnd(r) >> 32);
            if (y & 1) {
                const char* t = kw[zipf(r, 16)];
                p = put(d, p, n, t, strlen(t));
            } else {
                const uint32_t w = zipf(r, 1024), w2 = zipf(r, 256);
                p = put(d, p, n, v->w[w], v->len[w]);
                if (y & 2) {
                    p = put(d, p, n, "_", 1);
                    p = put(d, p, n, v->w[w2], v->len[w2]);
                }
                p = put(d, p, n, (y & 4) ? "(" : (y & 8) ? ", " : " ", (y & 4) ? 1 : (y & 8) ? 2 : 1);
            }
        }
        p = put(d, p, n, (x >> 8 & 3) ? ";\n" : ")\n{\n", (x >> 8 & 3) ? 2 : 4);
        if ((x >> 12 & 31) == 0 && p > 2048) { /* repeated boilerplate block */
            const size_t len = 40 + rndn(r, 400), back = 200 + rndn(r, p > 50000 ? 50000 : (uint32_t)p - 200);
            for (size_t k = 0; k < len && p < n; k++, p++) d[p] = d[p - back];
        }
    }
}

static void gen_float(uint8_t* d, size_t n, rng_t* r) {
    /* star-catalogue-like 28-byte records: mostly noise mantissas, a few structured bytes */
    size_t p = 0;
    uint32_t ra = 0;
    while (p < n) {
        uint8_t rec[28];
        for (int k = 0; k < 28; k += 8) {
            const uint64_t x = rnd(r);
            memcpy(rec + k, &x, k + 8 <= 28 ? 8 : 4);
        }
        ra += rndn(r, 5000);
        memcpy(rec, &ra, 4);
        rec[7] = 0x40;
        rec[15] = (uint8_t)(0x3F + rndn(r, 2));
        rec[26] = (uint8_t)rndn(r, 12);
        rec[27] = 0;
        p = put(d, p, n, rec, 28);
    }
}

static void gen_xml(uint8_t* d, size_t n, rng_t* r, const vocab_t* v) {
    static const char* tag[8] = {"record", "item", "name", "value", "entry", "field", "row", "node"};
    static const char* att[6] = {" id=\"", " type=\"", " lang=\"en\"", " xml:space=\"preserve\"", " ref=\"", " class=\""};
    size_t p = 0;
    int depth = 0;
    int stack[16];
    while (p < n) {
        const uint32_t x = (uint32_t)(rnd(r) >> 32);
        if (depth < 6 && ((x & 3) != 0 || depth == 0)) {
            const int t = (int)zipf(r, 8);
            for (int k = 0; k < depth; k++) p = put(d, p, n, "  ", 2);
            p = put(d, p, n, "<", 1);
            p = put(d, p, n, tag[t], strlen(tag[t]));
            for (unsigned a = 0; a < (x >> 2 & 3); a++) {
                const int ai = (int)rndn(r, 6);
                p = put(d, p, n, att[ai], strlen(att[ai]));
                if (att[ai][strlen(att[ai]) - 1] == '"' && (ai == 0 || ai == 1 || ai >= 4)) {
                    if (ai == 0 || ai == 4) p = put_num(d, p, n, zipf(r, 100000), 1);
                    else { const uint32_t w = zipf(r, 64); p = put(d, p, n, v->w[w], v->len[w]); }
                    p = put(d, p, n, "\"", 1);
                }
            }
            if ((x >> 6 & 3) == 0) {
                p = put(d, p, n, "/>\n", 3);
            } else if ((x >> 6 & 3) == 1) {
                p = put(d, p, n, ">", 1);
                const unsigned words = 1 + (x >> 8 & 7);
                for (unsigned k = 0; k < words; k++) {
                    const uint32_t w = zipf(r, VOCAB);
                    p = put(d, p, n, v->w[w], v->len[w]);
                    if (k + 1 < words) p = put(d, p, n, " ", 1);
                }
                p = put(d, p, n, "</", 2);
                p = put(d, p, n, tag[t], strlen(tag[t]));
                p = put(d, p, n, ">\n", 2);
            } else {
                p = put(d, p, n, ">\n", 2);
                stack[depth++] = t;
            }
        } else if (depth > 0) {
            const int t = stack[--depth];
            for (int k = 0; k < depth; k++) p = put(d, p, n, "  ", 2);
            p = put(d, p, n, "</", 2);
            p = put(d, p, n, tag[t], strlen(tag[t]));
            p = put(d, p, n, ">\n", 2);
        }
    }
}

static vocab_t g_vocab[3];
static int g_vocab_ready = 0;

static void vocab_init(void) {
    if (g_vocab_ready) return;
    vocab_build(&g_vocab[0], 0x1234567, "etaoinshrdlcumwfgypbvkjxqz", 26);
    vocab_build(&g_vocab[1], 0x7654321, "aieoznrwsctykdpmjulbghfxvq", 26);
    vocab_build(&g_vocab[2], 0xABCDEF1, "etaoinsrhldcumfpgwybvkxjqz", 26);
    g_vocab_ready = 1;
}

/* Call once from one thread before any parallel zxcorp_fill. */
void zxcorp_init(void) { vocab_init(); }

/* Fill dst[0..len) with corpus bytes [offset, offset+len); offset must be CHUNK-aligned. */
int zxcorp_fill(uint8_t* dst, uint64_t offset, uint64_t len, uint64_t seed) {
    if (offset % CHUNK) return -1;
    vocab_init();
    for (uint64_t done = 0; done < len; done += CHUNK) {
        const uint64_t c = (offset + done) / CHUNK;
        const size_t n = len - done < CHUNK ? (size_t)(len - done) : CHUNK;
        rng_t r = {(seed * 0x9E3779B97F4A7C15ull) ^ (c * 0xD1B54A32D192ED03ull) ^ 0x5851F42D4C957F2Dull};
        if (r.s == 0) r.s = 1;
        rnd(&r);
        uint8_t* d = dst + done;
        switch (unit_class((unsigned)(c % UNIT_CHUNKS))) {
            case K_TEXT_EN: gen_text(d, n, &r, &g_vocab[0], 0); break;
            case K_TEXT_PL: gen_text(d, n, &r, &g_vocab[1], 0); break;
            case K_TEXT_DICT: gen_text(d, n, &r, &g_vocab[2], 1); break;
            case K_EXE: gen_exe(d, n, &r, c / UNIT_CHUNKS); break;
            case K_IMG: gen_img(d, n, &r); break;
            case K_CHEM: gen_chem(d, n, &r); break;
            case K_DB: gen_db(d, n, &r, &g_vocab[0]); break;
            case K_SRC: gen_src(d, n, &r, &g_vocab[2]); break;
            case K_FLOAT: gen_float(d, n, &r); break;
            default: gen_xml(d, n, &r, &g_vocab[0]); break;
        }
    }
    return 0;
}

int zxcorp_class_of_chunk(uint64_t chunk) { return unit_class((unsigned)(chunk % UNIT_CHUNKS)); }

/* ------------------------------------------------------------------------- */
/* config 4 input: fixed-size JSON-ish records from a 200-key  */
/* schema with Zipf-distributed values; record r is a pure function of (seed,r).*/
/* ------------------------------------------------------------------------- */
int zxcorp_records(uint8_t* dst, uint64_t first_record, uint64_t n_records, uint32_t record_size, uint64_t seed) {
    vocab_init();
    const vocab_t* v = &g_vocab[0];
    for (uint64_t r = 0; r < n_records; r++) {
        rng_t g = {(seed * 0x9E3779B97F4A7C15ull) ^ ((first_record + r) * 0xD1B54A32D192ED03ull) ^ 0x1234ABCDull};
        if (g.s == 0) g.s = 1;
        rnd(&g);
        uint8_t* d = dst + r * record_size;
        size_t p = 0;
        p = put(d, p, record_size, "{", 1);
        int first = 1;
        while (p + 48 < record_size) {
            const uint32_t key = zipf(&g, 200);
            const uint32_t x = (uint32_t)(rnd(&g) >> 32);
            if (!first) p = put(d, p, record_size, ",", 1);
            first = 0;
            p = put(d, p, record_size, "\"", 1);
            p = put(d, p, record_size, v->w[key * 7 % VOCAB], v->len[key * 7 % VOCAB]);
            p = put(d, p, record_size, "_", 1);
            p = put(d, p, record_size, v->w[key], v->len[key]);
            p = put(d, p, record_size, "\":", 2);
            switch (key % 4) {
                case 0: p = put_num(d, p, record_size, zipf(&g, 1000000), 1); break;
                case 1: {
                    const uint32_t w = zipf(&g, 512);
                    p = put(d, p, record_size, "\"", 1);
                    p = put(d, p, record_size, v->w[w], v->len[w]);
                    p = put(d, p, record_size, "\"", 1);
                    break;
                }
                case 2: p = put(d, p, record_size, (x & 1) ? "true" : "false", (x & 1) ? 4 : 5); break;
                default: {
                    char t[32];
                    const int L = snprintf(t, sizeof t, "\"2026-%02u-%02uT%02u:%02u:%02uZ\"", 1 + x % 12, 1 + (x >> 4) % 28,
                                           (x >> 9) % 24, (x >> 14) % 60, (x >> 20) % 60);
                    p = put(d, p, record_size, t, (size_t)L);
                }
            }
        }
        p = put(d, p, record_size, "}", 1);
        while (p < record_size) { d[p] = (p + 1 == record_size) ? '\n' : ' '; p++; }
    }
    return 0;
}
