/* ORACLE (test infrastructure only) — Keccak-f[1600], SHA3-512, SHAKE256 (FIPS 202), keyed and
 * personalised BLAKE2b-512 (RFC 7693), STROBE-128 + Merlin 3.0.0 transcripts and TranscriptRng,
 * ChaCha12Rng (rand_chacha 0.3.1 + rand_core 0.6 seed_from_u64).
 * Call sites restated: /root/reference/src/protocols/curve_point_protocol.rs:31-35 (SHA3-512),
 * /root/reference/src/generators/generators_chain.rs:23-49 (SHAKE256),
 * /root/reference/src/utils/generic.rs:30-60 (BLAKE2b nonce),
 * /root/reference/src/transcripts.rs:59-194 and protocols/transcript_protocol.rs:39-79 (Merlin),
 * /root/reference/benches/range_proof.rs:47 and tests/ristretto.rs (ChaCha12Rng::seed_from_u64). */
#include "orc_internal.h"

/* ------------------------------------------------------------------ Keccak */
static const uint64_t RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int ROTC[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
static const int PILN[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
#define ROL64(x, n) (((x) << (n)) | ((x) >> (64 - (n))))

void keccak_f1600(uint64_t st[25]) {
    uint64_t bc[5], t;
    for (int r = 0; r < 24; r++) {
        for (int i = 0; i < 5; i++) bc[i] = st[i] ^ st[i + 5] ^ st[i + 10] ^ st[i + 15] ^ st[i + 20];
        for (int i = 0; i < 5; i++) {
            t = bc[(i + 4) % 5] ^ ROL64(bc[(i + 1) % 5], 1);
            for (int j = 0; j < 25; j += 5) st[j + i] ^= t;
        }
        t = st[1];
        for (int i = 0; i < 24; i++) {
            int j = PILN[i];
            bc[0] = st[j];
            st[j] = ROL64(t, ROTC[i]);
            t = bc[0];
        }
        for (int j = 0; j < 25; j += 5) {
            for (int i = 0; i < 5; i++) bc[i] = st[j + i];
            for (int i = 0; i < 5; i++) st[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
        }
        st[0] ^= RC[r];
    }
}

static void sponge_xor_byte(keccak_sponge *s, unsigned pos, uint8_t b) {
    s->st[pos / 8] ^= (uint64_t)b << (8 * (pos % 8));
}

static void sponge_absorb(keccak_sponge *s, const uint8_t *in, size_t len) {
    for (size_t i = 0; i < len; i++) {
        sponge_xor_byte(s, s->pos++, in[i]);
        if (s->pos == s->rate) { keccak_f1600(s->st); s->pos = 0; }
    }
}

static void sponge_pad(keccak_sponge *s, uint8_t dom) {
    sponge_xor_byte(s, s->pos, dom);
    sponge_xor_byte(s, s->rate - 1, 0x80);
    keccak_f1600(s->st);
    s->pos = 0;
}

static void sponge_squeeze(keccak_sponge *s, uint8_t *out, size_t len) {
    for (size_t i = 0; i < len; i++) {
        if (s->pos == s->rate) { keccak_f1600(s->st); s->pos = 0; }
        out[i] = (uint8_t)(s->st[s->pos / 8] >> (8 * (s->pos % 8)));
        s->pos++;
    }
}

void sha3_512(uint8_t out[64], const uint8_t *in, size_t len) {
    keccak_sponge s;
    memset(&s, 0, sizeof s);
    s.rate = 72;
    sponge_absorb(&s, in, len);
    sponge_pad(&s, 0x06);
    sponge_squeeze(&s, out, 64);
}

void shake256_init(keccak_sponge *s) { memset(s, 0, sizeof *s); s->rate = 136; }
void shake256_absorb(keccak_sponge *s, const uint8_t *in, size_t len) { sponge_absorb(s, in, len); }
void shake256_finalize(keccak_sponge *s) { sponge_pad(s, 0x1f); }
void shake256_squeeze(keccak_sponge *s, uint8_t *out, size_t len) { sponge_squeeze(s, out, len); }

/* ------------------------------------------------------------------ BLAKE2b */
static const uint64_t B2IV[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                                 0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
static const uint8_t B2SIGMA[12][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
#define ROR64(x, n) (((x) >> (n)) | ((x) << (64 - (n))))
#define B2G(a, b, c, d, x, y)          \
    do {                               \
        a = a + b + (x); d = ROR64(d ^ a, 32); \
        c = c + d;       b = ROR64(b ^ c, 24); \
        a = a + b + (y); d = ROR64(d ^ a, 16); \
        c = c + d;       b = ROR64(b ^ c, 63); \
    } while (0)

static void b2_compress(uint64_t h[8], const uint8_t block[128], u128 t, int last) {
    uint64_t m[16], v[16];
    for (int i = 0; i < 16; i++) {
        uint64_t x = 0;
        for (int j = 7; j >= 0; j--) x = (x << 8) | block[8 * i + j];
        m[i] = x;
    }
    for (int i = 0; i < 8; i++) { v[i] = h[i]; v[i + 8] = B2IV[i]; }
    v[12] ^= (uint64_t)t;
    v[13] ^= (uint64_t)(t >> 64);
    if (last) v[14] = ~v[14];
    for (int r = 0; r < 12; r++) {
        const uint8_t *s = B2SIGMA[r];
        B2G(v[0], v[4], v[8], v[12], m[s[0]], m[s[1]]);
        B2G(v[1], v[5], v[9], v[13], m[s[2]], m[s[3]]);
        B2G(v[2], v[6], v[10], v[14], m[s[4]], m[s[5]]);
        B2G(v[3], v[7], v[11], v[15], m[s[6]], m[s[7]]);
        B2G(v[0], v[5], v[10], v[15], m[s[8]], m[s[9]]);
        B2G(v[1], v[6], v[11], v[12], m[s[10]], m[s[11]]);
        B2G(v[2], v[7], v[8], v[13], m[s[12]], m[s[13]]);
        B2G(v[3], v[4], v[9], v[14], m[s[14]], m[s[15]]);
    }
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
}

void blake2b_keyed_personal_512(uint8_t out[64], const uint8_t *key, size_t keylen,
                                const uint8_t *person, size_t personlen,
                                const uint8_t *msg, size_t msglen) {
    uint8_t param[64];
    memset(param, 0, 64);
    param[0] = 64;
    param[1] = (uint8_t)keylen;
    param[2] = 1;
    param[3] = 1;
    if (personlen > 16) personlen = 16;
    memcpy(param + 48, person, personlen);
    uint64_t h[8];
    for (int i = 0; i < 8; i++) {
        uint64_t x = 0;
        for (int j = 7; j >= 0; j--) x = (x << 8) | param[8 * i + j];
        h[i] = B2IV[i] ^ x;
    }
    uint8_t block[128];
    u128 t = 0;
    if (keylen) {
        memset(block, 0, 128);
        memcpy(block, key, keylen);
        t = 128;
        if (msglen == 0) { b2_compress(h, block, t, 1); goto done; }
        b2_compress(h, block, t, 0);
    }
    while (msglen > 128) {
        t += 128;
        b2_compress(h, msg, t, 0);
        msg += 128; msglen -= 128;
    }
    memset(block, 0, 128);
    if (msglen) memcpy(block, msg, msglen);
    t += msglen;
    b2_compress(h, block, t, 1);
done:
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) out[8 * i + j] = (uint8_t)(h[i] >> (8 * j));
}

/* ------------------------------------------------------------------ STROBE-128 / Merlin */
#define STROBE_R 166
#define FLAG_I 1
#define FLAG_A 2
#define FLAG_C 4
#define FLAG_T 8
#define FLAG_M 16
#define FLAG_K 32

static void strobe_permute(strobe128 *s) {
    uint64_t w[25];
    for (int i = 0; i < 25; i++) {
        uint64_t x = 0;
        for (int j = 7; j >= 0; j--) x = (x << 8) | s->st[8 * i + j];
        w[i] = x;
    }
    keccak_f1600(w);
    for (int i = 0; i < 25; i++)
        for (int j = 0; j < 8; j++) s->st[8 * i + j] = (uint8_t)(w[i] >> (8 * j));
}

static void strobe_run_f(strobe128 *s) {
    s->st[s->pos] ^= s->pos_begin;
    s->st[s->pos + 1] ^= 0x04;
    s->st[STROBE_R + 1] ^= 0x80;
    strobe_permute(s);
    s->pos = 0;
    s->pos_begin = 0;
}

static void strobe_absorb(strobe128 *s, const uint8_t *d, size_t len) {
    for (size_t i = 0; i < len; i++) {
        s->st[s->pos] ^= d[i];
        s->pos++;
        if (s->pos == STROBE_R) strobe_run_f(s);
    }
}

static void strobe_overwrite(strobe128 *s, const uint8_t *d, size_t len) {
    for (size_t i = 0; i < len; i++) {
        s->st[s->pos] = d[i];
        s->pos++;
        if (s->pos == STROBE_R) strobe_run_f(s);
    }
}

static void strobe_squeeze(strobe128 *s, uint8_t *d, size_t len) {
    for (size_t i = 0; i < len; i++) {
        d[i] = s->st[s->pos];
        s->st[s->pos] = 0;
        s->pos++;
        if (s->pos == STROBE_R) strobe_run_f(s);
    }
}

static void strobe_begin_op(strobe128 *s, uint8_t flags, int more) {
    if (more) return; /* caller guarantees same flags */
    uint8_t old_begin = s->pos_begin;
    s->pos_begin = (uint8_t)(s->pos + 1);
    s->cur_flags = flags;
    uint8_t hdr[2] = {old_begin, flags};
    strobe_absorb(s, hdr, 2);
    int force_f = (flags & (FLAG_C | FLAG_K)) != 0;
    if (force_f && s->pos != 0) strobe_run_f(s);
}

static void strobe_meta_ad(strobe128 *s, const uint8_t *d, size_t len, int more) {
    strobe_begin_op(s, FLAG_M | FLAG_A, more);
    strobe_absorb(s, d, len);
}
static void strobe_ad(strobe128 *s, const uint8_t *d, size_t len, int more) {
    strobe_begin_op(s, FLAG_A, more);
    strobe_absorb(s, d, len);
}
static void strobe_prf(strobe128 *s, uint8_t *d, size_t len, int more) {
    strobe_begin_op(s, FLAG_I | FLAG_A | FLAG_C, more);
    strobe_squeeze(s, d, len);
}
static void strobe_key(strobe128 *s, const uint8_t *d, size_t len, int more) {
    strobe_begin_op(s, FLAG_A | FLAG_C, more);
    strobe_overwrite(s, d, len);
}

static void strobe_new(strobe128 *s, const uint8_t *proto, size_t len) {
    memset(s, 0, sizeof *s);
    static const uint8_t hdr[6] = {1, STROBE_R + 2, 1, 0, 1, 96};
    memcpy(s->st, hdr, 6);
    memcpy(s->st + 6, "STROBEv1.0.2", 12);
    strobe_permute(s);
    s->pos = 0; s->pos_begin = 0; s->cur_flags = 0;
    strobe_meta_ad(s, proto, len, 0);
}

static void le32(uint8_t b[4], uint32_t x) {
    b[0] = (uint8_t)x; b[1] = (uint8_t)(x >> 8); b[2] = (uint8_t)(x >> 16); b[3] = (uint8_t)(x >> 24);
}

void merlin_init(merlin_transcript *t, const uint8_t *label, size_t len) {
    strobe_new(&t->s, (const uint8_t *)"Merlin v1.0", 11);
    merlin_append_message(t, "dom-sep", label, len);
}

void merlin_append_message(merlin_transcript *t, const char *label, const uint8_t *msg, size_t len) {
    uint8_t l4[4];
    le32(l4, (uint32_t)len);
    strobe_meta_ad(&t->s, (const uint8_t *)label, strlen(label), 0);
    strobe_meta_ad(&t->s, l4, 4, 1);
    strobe_ad(&t->s, msg, len, 0);
}

void merlin_append_u64(merlin_transcript *t, const char *label, uint64_t x) {
    uint8_t b[8];
    for (int i = 0; i < 8; i++) b[i] = (uint8_t)(x >> (8 * i));
    merlin_append_message(t, label, b, 8);
}

void merlin_challenge_bytes(merlin_transcript *t, const char *label, uint8_t *out, size_t len) {
    uint8_t l4[4];
    le32(l4, (uint32_t)len);
    strobe_meta_ad(&t->s, (const uint8_t *)label, strlen(label), 0);
    strobe_meta_ad(&t->s, l4, 4, 1);
    strobe_prf(&t->s, out, len, 0);
}

void merlin_build_rng(merlin_rng *r, const merlin_transcript *t, const uint8_t *witness, size_t wlen,
                      int have_witness, orc_rng *ext) {
    r->s = t->s;
    if (have_witness) {
        uint8_t l4[4];
        le32(l4, (uint32_t)wlen);
        strobe_meta_ad(&r->s, (const uint8_t *)"witness", 7, 0);
        strobe_meta_ad(&r->s, l4, 4, 1);
        strobe_key(&r->s, witness, wlen, 0);
    }
    uint8_t rb[32];
    ext->fill(ext, rb, 32);
    strobe_meta_ad(&r->s, (const uint8_t *)"rng", 3, 0);
    strobe_key(&r->s, rb, 32, 0);
}

void merlin_rng_fill(merlin_rng *r, uint8_t *dst, size_t len) {
    uint8_t l4[4];
    le32(l4, (uint32_t)len);
    strobe_meta_ad(&r->s, l4, 4, 0);
    strobe_prf(&r->s, dst, len, 0);
}

/* ------------------------------------------------------------------ RNGs */
#define ROL32(x, n) (((x) << (n)) | ((x) >> (32 - (n))))
#define QR(a, b, c, d)                 \
    a += b; d ^= a; d = ROL32(d, 16);  \
    c += d; b ^= c; b = ROL32(b, 12);  \
    a += b; d ^= a; d = ROL32(d, 8);   \
    c += d; b ^= c; b = ROL32(b, 7);

static void chacha12_block(uint32_t out[16], const uint32_t key[8], uint64_t counter) {
    uint32_t x[16] = {0x61707865, 0x3320646e, 0x79622d32, 0x6b206574,
                      key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                      (uint32_t)counter, (uint32_t)(counter >> 32), 0, 0};
    uint32_t in[16];
    memcpy(in, x, sizeof x);
    for (int i = 0; i < 6; i++) {
        QR(x[0], x[4], x[8], x[12]) QR(x[1], x[5], x[9], x[13]) QR(x[2], x[6], x[10], x[14]) QR(x[3], x[7], x[11], x[15])
        QR(x[0], x[5], x[10], x[15]) QR(x[1], x[6], x[11], x[12]) QR(x[2], x[7], x[8], x[13]) QR(x[3], x[4], x[9], x[14])
    }
    for (int i = 0; i < 16; i++) out[i] = x[i] + in[i];
}

static void chacha12_refill(chacha12_rng *r) {
    for (int b = 0; b < 4; b++) chacha12_block(r->buf + 16 * b, r->key, r->counter + (uint64_t)b);
    r->counter += 4;
}

static void chacha12_fill(orc_rng *self, uint8_t *dst, size_t len) {
    chacha12_rng *r = (chacha12_rng *)self;
    size_t done = 0;
    while (done < len) {
        if (r->index >= 64) { chacha12_refill(r); r->index = 0; }
        /* rand_core fill_via_u32_chunks: whole words consumed, a trailing partial word is discarded */
        size_t avail_words = 64 - r->index;
        size_t want = len - done;
        size_t words = (want + 3) / 4;
        if (words > avail_words) words = avail_words;
        size_t bytes = words * 4 < want ? words * 4 : want;
        for (size_t i = 0; i < bytes; i++) dst[done + i] = (uint8_t)(r->buf[r->index + i / 4] >> (8 * (i % 4)));
        r->index += (unsigned)words;
        done += bytes;
    }
}

void chacha12_from_seed(chacha12_rng *r, const uint8_t seed[32]) {
    memset(r, 0, sizeof *r);
    r->base.fill = chacha12_fill;
    for (int i = 0; i < 8; i++)
        r->key[i] = (uint32_t)seed[4 * i] | ((uint32_t)seed[4 * i + 1] << 8) | ((uint32_t)seed[4 * i + 2] << 16) | ((uint32_t)seed[4 * i + 3] << 24);
    r->counter = 0;
    r->index = 64;
}

void chacha12_seed_from_u64(chacha12_rng *r, uint64_t state) {
    /* rand_core 0.6 SeedableRng::seed_from_u64: PCG32 stream fills the seed 4 bytes at a time */
    uint8_t seed[32];
    for (int i = 0; i < 8; i++) {
        state = state * 6364136223846793005ULL + 11634580027462260723ULL;
        uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
        uint32_t rot = (uint32_t)(state >> 59);
        uint32_t x = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
        le32(seed + 4 * i, x);
    }
    chacha12_from_seed(r, seed);
}

uint32_t chacha12_next_u32(chacha12_rng *r) {
    if (r->index >= 64) { chacha12_refill(r); r->index = 0; }
    return r->buf[r->index++];
}

uint64_t chacha12_next_u64(chacha12_rng *r) {
    /* rand_core BlockRng::next_u64 */
    if (r->index < 63) {
        uint64_t lo = r->buf[r->index], hi = r->buf[r->index + 1];
        r->index += 2;
        return (hi << 32) | lo;
    } else if (r->index >= 64) {
        chacha12_refill(r);
        r->index = 2;
        return ((uint64_t)r->buf[1] << 32) | r->buf[0];
    } else {
        uint64_t lo = r->buf[63];
        chacha12_refill(r);
        r->index = 1;
        return ((uint64_t)r->buf[0] << 32) | lo;
    }
}

static void null_fill(orc_rng *self, uint8_t *dst, size_t len) { (void)self; memset(dst, 0, len); }
void null_rng_init(null_rng *r) { r->base.fill = null_fill; }

static void buf_fill(orc_rng *self, uint8_t *dst, size_t len) {
    buf_rng *r = (buf_rng *)self;
    if (r->off + len > r->len) { r->underflow = 1; memset(dst, 0, len); return; }
    memcpy(dst, r->p + r->off, len);
    r->off += len;
}
void buf_rng_init(buf_rng *r, const uint8_t *p, size_t len) {
    r->base.fill = buf_fill; r->p = p; r->len = len; r->off = 0; r->underflow = 0;
}
