// placeholder until the whole-proof driver lands (replaced in the next milestone)
#include "ctx.h"
extern "C" int32_t p2g_circuit_load(p2g_ctx*, const p2g_circuit_desc*, p2g_circuit**, uint64_t*) { return P2G_E_BADARG; }
extern "C" int32_t p2g_circuit_free(p2g_ctx*, p2g_circuit*) { return P2G_E_BADARG; }
extern "C" size_t p2g_proof_words(const p2g_circuit*) { return 0; }
extern "C" int32_t p2g_prove(p2g_ctx*, const p2g_circuit*, const uint64_t*, const uint64_t*, uint64_t*, size_t, size_t*) { return P2G_E_BADARG; }
extern "C" int32_t p2g_prove_dev(p2g_ctx*, const p2g_circuit*, const uint64_t*, const uint64_t*, uint64_t*, size_t, size_t*) { return P2G_E_BADARG; }
extern "C" int32_t p2g_last_transcript(p2g_ctx*, p2g_transcript*) { return P2G_E_BADARG; }
extern "C" int32_t p2g_last_zs_values(p2g_ctx*, uint64_t*) { return P2G_E_BADARG; }
extern "C" int32_t p2g_last_quotient_chunks(p2g_ctx*, uint64_t*) { return P2G_E_BADARG; }
extern "C" int32_t p2g_last_timings(p2g_ctx*, p2g_timings*) { return P2G_E_BADARG; }
extern "C" int32_t p2g_set_timing(p2g_ctx*, int32_t) { return P2G_E_BADARG; }
extern "C" int32_t p2g_pow_grind(p2g_ctx*, const uint64_t*, uint32_t, uint32_t, uint64_t*) { return P2G_E_BADARG; }
extern "C" int32_t p2g_fri_fold(p2g_ctx*, const uint64_t*, uint32_t, uint32_t, uint64_t, const uint64_t*, uint64_t*) { return P2G_E_BADARG; }
