// sla_batch.cuh -- batch of independent instances, one CTA per instance (BASELINE.json config 4).
#pragma once

extern "C" {

void sla_batch_free(sla_ctx* ctx) { (void)ctx; }

int sla_batch_upload(sla_ctx* ctx, uint32_t, const uint32_t*, const uint32_t*, const uint32_t*, const uint32_t*,
                     const double*) {
    return fail(ctx, SLA_ERR_STATE, "batch engine not built yet");
}
int sla_batch_generate_device(sla_ctx* ctx, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint64_t, uint32_t, uint32_t,
                              int) {
    return fail(ctx, SLA_ERR_STATE, "batch engine not built yet");
}
int sla_batch_solve(sla_ctx* ctx, int, int, double, double, uint32_t, uint32_t*, uint32_t*, double*, sla_stats*,
                    sla_stats*) {
    return fail(ctx, SLA_ERR_STATE, "batch engine not built yet");
}

}  // extern "C"
