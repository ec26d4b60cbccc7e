// Entry points not implemented yet return GPAR_ERR_INVALID with a message (never a CPU fallback).
#include "common.cuh"
#define NOTYET(name) return ctx ? gpar_fail(ctx, GPAR_ERR_INVALID, name " is not implemented yet") : GPAR_ERR_INVALID
extern "C" {
int gpar_exact_logpdf(gpar_ctx* ctx, int, int, const double*, int32_t, double*) { NOTYET("gpar_exact_logpdf"); }
int gpar_exact_posterior(gpar_ctx* ctx, int, int, const double*, int32_t, const double*, int64_t, double*, double*) { NOTYET("gpar_exact_posterior"); }
}
