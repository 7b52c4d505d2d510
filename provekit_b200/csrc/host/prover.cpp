// provekit_b200/csrc/host/prover.cpp — placeholder until the host driver lands (next commit).
#include "../pk_internal.h"
struct pk_prover { pk_ctx* ctx; };
extern "C" {
int pk_prover_create(pk_ctx* ctx, const pk_r1cs*, pk_prover**) { return pk::set_err(ctx, PK_ERR_INTERNAL, "pk_prove: not built yet"); }
void pk_prover_destroy(pk_prover*) {}
int pk_prove(pk_prover*, const uint64_t*, const pk_rand*, uint8_t**, size_t*) { return PK_ERR_INTERNAL; }
void pk_free(void* p) { free(p); }
void pk_prover_timings(const pk_prover*, double out[9]) { for (int i = 0; i < 9; i++) out[i] = 0; }
}
