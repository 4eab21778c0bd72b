// synth.h -- counter-based synthetic instance generator shared by the host (sla_generate_host) and the device
// (sla_generate_device) so that both produce bit-identical CSR data (SURVEY.md 8d "Synthetic inputs").
//
// Shapes follow the reference's bench generators: k distinct sorted columns per row drawn as a uniform
// k-subset (benches/benchmark.rs:63-67), integer-valued f64 costs (benchmark.rs:73), and for the symmetric
// family one planted permutation arc per row so that a perfect matching exists (benchmark.rs:32-33,39).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define SLA_HD __host__ __device__ __forceinline__
#else
#define SLA_HD inline
#endif

namespace sla_synth {

SLA_HD uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

SLA_HD uint64_t draw(uint64_t seed, uint64_t stream, uint64_t row, uint64_t slot) {
    uint64_t h = splitmix64(seed * 0xD1342543DE82EF95ull + stream);
    h = splitmix64(h ^ (row * 0x9E3779B97F4A7C15ull));
    return splitmix64(h + slot);
}

// uniform integer in [0, n) by the multiply-high map (n < 2^32 keeps the bias below 2^-32)
SLA_HD uint32_t below(uint64_t h, uint32_t n) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__umul64hi(h, (uint64_t)n);
#else
    return (uint32_t)(((unsigned __int128)h * (unsigned __int128)n) >> 64);
#endif
}

struct Spec {
    uint32_t num_rows, num_cols, k;
    uint64_t seed;
    uint32_t value_lo, value_hi;
    uint32_t planted;        // 0 / 1
    uint32_t perm_a, perm_b; // planted column of row i = (perm_a * i + perm_b) mod num_rows
    uint32_t value_dist;     // 0: uniform integers in [value_lo, value_hi); 1: floor((hi - lo) * Beta(3,3) + lo), the
                             // reference's asymmetric bench values floor(700 * Beta(3,3) + 300) (benches/benchmark.rs:60,73)
};

// Uniform double in [0, 1) from the top 53 bits.
SLA_HD double unit(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }

// Beta(3,3) variate: the median (3rd order statistic) of five independent uniforms is exactly Beta(3, 3).
SLA_HD double beta33(uint64_t seed, uint64_t row, uint64_t slot) {
    double u[5];
    for (int t = 0; t < 5; ++t) u[t] = unit(draw(seed, 3 + (uint64_t)t, row, slot));
    // partial selection: two passes of a bubble pass leave the two largest at the end; the max of the rest is the median
    for (int pass = 0; pass < 2; ++pass)
        for (int t = 0; t + 1 < 5 - pass; ++t)
            if (u[t] > u[t + 1]) { const double x = u[t]; u[t] = u[t + 1]; u[t + 1] = x; }
    double m = u[0];
    if (u[1] > m) m = u[1];
    if (u[2] > m) m = u[2];
    return m;
}

SLA_HD uint32_t planted_col(const Spec& s, uint32_t row) {
    return (uint32_t)(((uint64_t)s.perm_a * row + s.perm_b) % s.num_rows);
}

// Writes the k sorted distinct columns and k values of `row` to out_cols / out_vals (row-local pointers).
SLA_HD void make_row(const Spec& s, uint32_t row, uint32_t* out_cols, double* out_vals) {
    const uint32_t k = s.k;
    uint32_t n_free = k, universe = s.num_cols, pcol = 0;
    if (s.planted) { n_free = k - 1; universe = s.num_cols - 1; pcol = planted_col(s, row); }
    // Floyd's algorithm: uniform n_free-subset of [0, universe)
    uint32_t cnt = 0;
    for (uint32_t t = 0; t < n_free; ++t) {
        const uint32_t j = universe - n_free + t;
        uint32_t r = below(draw(s.seed, 1, row, t), j + 1);
        bool seen = false;
        for (uint32_t u = 0; u < cnt; ++u) seen |= (out_cols[u] == r);
        out_cols[cnt++] = seen ? j : r;
    }
    if (s.planted) {
        for (uint32_t u = 0; u < cnt; ++u) out_cols[u] += (out_cols[u] >= pcol) ? 1u : 0u;
        out_cols[cnt++] = pcol;
    }
    // insertion sort (k is small)
    for (uint32_t a = 1; a < k; ++a) {
        uint32_t x = out_cols[a];
        uint32_t b = a;
        while (b > 0 && out_cols[b - 1] > x) { out_cols[b] = out_cols[b - 1]; --b; }
        out_cols[b] = x;
    }
    const uint32_t span = s.value_hi - s.value_lo;
    for (uint32_t t = 0; t < k; ++t) {
        if (s.value_dist == 1u) {
            uint32_t x = (uint32_t)((double)span * beta33(s.seed, row, t));
            if (x >= span) x = span - 1u;
            out_vals[t] = (double)(s.value_lo + x);
        } else {
            out_vals[t] = (double)(s.value_lo + below(draw(s.seed, 2, row, t), span));
        }
    }
}

SLA_HD uint32_t gcd_u32(uint32_t a, uint32_t b) { while (b) { uint32_t t = a % b; a = b; b = t; } return a; }

// Fills perm_a / perm_b from the seed (perm_a coprime to num_rows => a bijection on [0, num_rows)).
SLA_HD void finish_spec(Spec& s) {
    s.perm_a = 1; s.perm_b = 0;
    if (!s.planted || s.num_rows == 0) return;
    uint64_t h = splitmix64(s.seed ^ 0xA5A5A5A5DEADBEEFull);
    uint32_t a = (uint32_t)(h % s.num_rows);
    if (a == 0) a = 1;
    while (gcd_u32(a, s.num_rows) != 1) { a += 1; if (a >= s.num_rows) a = 1; }
    s.perm_a = a;
    s.perm_b = (uint32_t)(splitmix64(h) % s.num_rows);
}

}  // namespace sla_synth
