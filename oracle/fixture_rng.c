/*
 * fixture_rng.c -- TEST INFRASTRUCTURE ONLY (part of the CPU oracle, never on the product path).
 *
 * Regenerates, bit for bit, the random k-sparse fixtures of the reference's own tests
 * (`populate_with_ksparse_input`, /root/reference/src/solver.rs:261-292) without a Rust toolchain.
 * The third-party crates involved are dev-dependencies of the reference (Cargo.toml:24-27, semver ranges,
 * no Cargo.lock): rand 0.8, rand_chacha 0.3, reservoir-sampling 0.5.  Their published algorithms are
 * restated here:
 *   - SeedableRng::seed_from_u64  : PCG32 expansion of the u64 into a 32-byte seed (rand_core 0.6)
 *   - ChaCha8Rng                  : ChaCha, 8 rounds, 64-bit block counter, 4 blocks buffered (rand_chacha 0.3)
 *   - Uniform<f64>::sample        : 52 random mantissa bits -> [1,2) - 1, scaled (rand 0.8)
 *   - Rng::gen_range(0..n) usize  : widening-multiply rejection sampling with the "zone" test (rand 0.8)
 *   - reservoir_sampling::unweighted::core::r : algorithm R with an EXCLUSIVE upper bound on the draw
 * The chain is pinned by the golden objectives at solver.rs:296, 332-336, 435 (tests/test_oracle_goldens.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint32_t key[8];
    uint64_t counter;   /* block counter (words 12-13) */
    uint32_t buf[64];   /* 4 blocks */
    uint32_t index;     /* next unread word; 64 = empty */
} chacha8;

static uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
static uint32_t rotr32(uint32_t x, uint32_t n) { n &= 31; return n ? ((x >> n) | (x << (32 - n))) : x; }

#define QR(a, b, c, d)                                                    \
    a += b; d ^= a; d = rotl32(d, 16); c += d; b ^= c; b = rotl32(b, 12); \
    a += b; d ^= a; d = rotl32(d, 8);  c += d; b ^= c; b = rotl32(b, 7);

static void chacha8_block(const uint32_t key[8], uint64_t counter, uint32_t out[16]) {
    uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
    for (int i = 0; i < 8; ++i) in[4 + i] = key[i];
    in[12] = (uint32_t)counter;
    in[13] = (uint32_t)(counter >> 32);
    in[14] = 0; in[15] = 0; /* stream id */
    uint32_t x[16];
    memcpy(x, in, sizeof x);
    for (int r = 0; r < 4; ++r) { /* 8 rounds = 4 double rounds */
        QR(x[0], x[4], x[8], x[12]) QR(x[1], x[5], x[9], x[13]) QR(x[2], x[6], x[10], x[14]) QR(x[3], x[7], x[11], x[15])
        QR(x[0], x[5], x[10], x[15]) QR(x[1], x[6], x[11], x[12]) QR(x[2], x[7], x[8], x[13]) QR(x[3], x[4], x[9], x[14])
    }
    for (int i = 0; i < 16; ++i) out[i] = x[i] + in[i];
}

static void chacha8_refill(chacha8 *g) {
    for (int b = 0; b < 4; ++b) chacha8_block(g->key, g->counter + (uint64_t)b, g->buf + 16 * b);
    g->counter += 4;
}

static void chacha8_seed_from_u64(chacha8 *g, uint64_t state) {
    const uint64_t MUL = 6364136223846793005ULL, INC = 11634580027462260723ULL;
    for (int i = 0; i < 8; ++i) {
        state = state * MUL + INC;
        uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
        uint32_t rot = (uint32_t)(state >> 59);
        g->key[i] = rotr32(xorshifted, rot);
    }
    g->counter = 0;
    g->index = 64;
}

/* BlockRng::next_u64 (rand_core 0.6 block.rs) */
static uint64_t chacha8_next_u64(chacha8 *g) {
    uint32_t idx = g->index;
    if (idx < 63) {
        g->index = idx + 2;
        return (uint64_t)g->buf[idx] | ((uint64_t)g->buf[idx + 1] << 32);
    } else if (idx >= 64) {
        chacha8_refill(g);
        g->index = 2;
        return (uint64_t)g->buf[0] | ((uint64_t)g->buf[1] << 32);
    } else {
        uint64_t lo = g->buf[63];
        chacha8_refill(g);
        g->index = 1;
        return lo | ((uint64_t)g->buf[0] << 32);
    }
}

/* Uniform::<f64>::from(low..high).sample(rng) */
static double uniform_f64(chacha8 *g, double low, double high) {
    double scale = high - low;
    uint64_t bits = (chacha8_next_u64(g) >> 12) | 0x3FF0000000000000ULL;
    double value1_2;
    memcpy(&value1_2, &bits, sizeof bits);
    return (value1_2 - 1.0) * scale + low;
}

/* rng.gen_range(0..range) for usize on a 64-bit target */
static uint64_t gen_range_u64(chacha8 *g, uint64_t range) {
    uint64_t zone = (range << __builtin_clzll(range)) - 1;
    for (;;) {
        uint64_t v = chacha8_next_u64(g);
        unsigned __int128 m = (unsigned __int128)v * range;
        uint64_t hi = (uint64_t)(m >> 64), lo = (uint64_t)m;
        if (lo <= zone) return hi;
    }
}

static int cmp_u32(const void *a, const void *b) {
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return (x > y) - (x < y);
}

/*
 * populate_with_ksparse_input (solver.rs:261-292): val_rng seed `val_seed` (1 in the tests), filter_rng seed
 * `filter_seed` (2).  Per row: reservoir-sample k columns from 0..num_cols with filter_rng, sort them, then draw
 * k values U(0, max_value) from val_rng.  Writes a CSR (row_ptr has num_rows+1 entries).
 */
void orc_fixture_ksparse(uint64_t val_seed, uint64_t filter_seed, uint32_t num_rows, uint32_t num_cols, uint32_t k,
                         double max_value, uint32_t *row_ptr, uint32_t *cols, double *vals) {
    chacha8 val_rng, filter_rng;
    chacha8_seed_from_u64(&val_rng, val_seed);
    chacha8_seed_from_u64(&filter_rng, filter_seed);
    row_ptr[0] = 0;
    for (uint32_t i = 0; i < num_rows; ++i) {
        uint32_t *slot = cols + (size_t)i * k;
        for (uint32_t t = 0; t < k; ++t) slot[t] = 0; /* vec![0; k] */
        for (uint32_t item = 0; item < num_cols; ++item) {
            if (item < k) {
                slot[item] = item;
            } else {
                uint64_t j = gen_range_u64(&filter_rng, item);
                if (j < k) slot[j] = item;
            }
        }
        qsort(slot, k, sizeof(uint32_t), cmp_u32);
        for (uint32_t t = 0; t < k; ++t) vals[(size_t)i * k + t] = uniform_f64(&val_rng, 0.0, max_value);
        row_ptr[i + 1] = (i + 1) * k;
    }
}

/* Raw stream access, for unit-testing the generator itself. */
void orc_chacha8_u64_stream(uint64_t seed, uint64_t *out, size_t n) {
    chacha8 g;
    chacha8_seed_from_u64(&g, seed);
    for (size_t i = 0; i < n; ++i) out[i] = chacha8_next_u64(&g);
}
