// sla_host_impl.h -- host-only entry points of include/sla.h (no CUDA): the synthetic instance generator.
// Compiled twice: into libsla_b200.so (sla_api.cu) and, with plain g++, into libsla_host.so (sla_host.cpp), so that
// input generation for a CPU-only process -- bench.py's reference arm, the oracle tests -- maps no CUDA library.
#pragma once
#include <cstdint>
#include <thread>
#include <vector>

#include "../../include/sla.h"
#include "synth.h"

namespace sla_hostgen {

inline int make_spec(sla_synth::Spec* s, uint32_t num_rows, uint32_t num_cols, uint32_t k, uint64_t seed, uint32_t value_lo,
                     uint32_t value_hi, int planted, int value_dist) {
    if (k == 0 || k > num_cols || (planted && num_cols < 2 && k > 1)) return SLA_ERR_INVALID;
    if (!(value_hi > value_lo)) return SLA_ERR_INVALID;
    if ((uint64_t)num_rows * k >= 0xFFFFFFFFull) return SLA_ERR_INVALID;
    if (value_dist != 0 && value_dist != 1) return SLA_ERR_INVALID;
    s->num_rows = num_rows; s->num_cols = num_cols; s->k = k; s->seed = seed;
    s->value_lo = value_lo; s->value_hi = value_hi; s->planted = planted ? 1u : 0u;
    s->value_dist = (uint32_t)value_dist;
    sla_synth::finish_spec(*s);
    return SLA_OK;
}

// Rows [row_begin, row_begin + row_count) of the global instance; row_ptr is local to the shard (row_count + 1 entries).
inline int generate_rows(const sla_synth::Spec& s, uint32_t row_begin, uint32_t row_count, int threads, uint32_t* row_ptr,
                         uint32_t* column_indices, double* values) {
    if (!row_ptr || !column_indices || !values) return SLA_ERR_INVALID;
    if ((uint64_t)row_begin + row_count > s.num_rows) return SLA_ERR_INVALID;
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    if (row_count < (1u << 14)) threads = 1;
    auto work = [&](uint32_t lo, uint32_t hi) {
        for (uint32_t r = lo; r < hi; ++r) {
            const size_t off = (size_t)r * s.k;
            sla_synth::make_row(s, row_begin + r, column_indices + off, values + off);
            row_ptr[r] = (uint32_t)off;
        }
    };
    std::vector<std::thread> pool;
    const uint32_t chunk = (row_count + (uint32_t)threads - 1) / (uint32_t)threads;
    for (int t = 1; t < threads; ++t) {
        const uint32_t lo = (uint32_t)t * chunk, hi = lo + chunk < row_count ? lo + chunk : row_count;
        if (lo < hi) pool.emplace_back(work, lo, hi);
    }
    work(0, chunk < row_count ? chunk : row_count);
    for (auto& th : pool) th.join();
    row_ptr[row_count] = row_count * s.k;
    return SLA_OK;
}

}  // namespace sla_hostgen

extern "C" {

int sla_generate_host(uint32_t num_rows, uint32_t num_cols, uint32_t k, uint64_t seed, uint32_t value_lo,
                      uint32_t value_hi, int planted, uint32_t* row_ptr, uint32_t* column_indices, double* values) {
    sla_synth::Spec s;
    int rc = sla_hostgen::make_spec(&s, num_rows, num_cols, k, seed, value_lo, value_hi, planted, 0);
    if (rc) return rc;
    return sla_hostgen::generate_rows(s, 0, num_rows, 1, row_ptr, column_indices, values);
}

int sla_generate_host_ex(uint32_t global_rows, uint32_t num_cols, uint32_t k, uint64_t seed, uint32_t value_lo,
                         uint32_t value_hi, int planted, int value_dist, uint32_t row_begin, uint32_t row_count, int threads,
                         uint32_t* row_ptr, uint32_t* column_indices, double* values) {
    sla_synth::Spec s;
    int rc = sla_hostgen::make_spec(&s, global_rows, num_cols, k, seed, value_lo, value_hi, planted, value_dist);
    if (rc) return rc;
    return sla_hostgen::generate_rows(s, row_begin, row_count, threads, row_ptr, column_indices, values);
}

}  // extern "C"
