// sla_part.cuh -- row-partitioned single instance across ranks (BASELINE.json config 5).
#pragma once

extern "C" {

void sla_part_free(sla_ctx* ctx) { (void)ctx; }

int sla_part_begin(sla_ctx* ctx, int, int, uint32_t, uint32_t, double, double, double) {
    return fail(ctx, SLA_ERR_STATE, "partitioned engine not built yet");
}
int sla_part_local_value_range(sla_ctx* ctx, double*, double*, double*) {
    return fail(ctx, SLA_ERR_STATE, "partitioned engine not built yet");
}
int sla_part_bid(sla_ctx* ctx) { return fail(ctx, SLA_ERR_STATE, "partitioned engine not built yet"); }
int sla_part_buffers(sla_ctx* ctx, void**, void**, uint64_t*) {
    return fail(ctx, SLA_ERR_STATE, "partitioned engine not built yet");
}
int sla_part_assign(sla_ctx* ctx, uint32_t*, uint32_t*) {
    return fail(ctx, SLA_ERR_STATE, "partitioned engine not built yet");
}
int sla_part_finish(sla_ctx* ctx, uint32_t*, uint32_t*, double*, sla_stats*) {
    return fail(ctx, SLA_ERR_STATE, "partitioned engine not built yet");
}

}  // extern "C"
