// Entry points declared in include/mc_b200.h whose kernels are still being written.  Each fails loudly.
#include "mc_common.cuh"
#define TODO(name) { mc_set_error(name ": not implemented yet in this build"); return MC_ERR_UNSUPPORTED; }
extern "C" int mc_env_find(mc_ctx*, const mc_table*, const char* const*, const char* const*, int, int, int, int64_t, int64_t, int, int, const char*, int*) TODO("mc_env_find")
extern "C" int mc_bfs(mc_ctx*, const mc_table*, const char*, int, int, int64_t, int64_t, int, uint64_t*, int32_t*, int16_t*, size_t, size_t*) TODO("mc_bfs")
extern "C" int mc_env_kmers(mc_ctx*, const mc_table*, const char*, int, int, int64_t, int64_t, int, uint64_t*, int32_t*, size_t, size_t*) TODO("mc_env_kmers")
extern "C" int mc_targets_load(const char*, mc_targets**) TODO("mc_targets_load")
extern "C" int mc_targets_count(const mc_targets*) { return 0; }
extern "C" const char* mc_targets_seq(const mc_targets*, int) { return ""; }
extern "C" const char* mc_targets_name(const mc_targets*, int) { return ""; }
extern "C" void mc_targets_free(mc_targets*) {}
extern "C" int mc_env_multi(mc_ctx*, const char* const*, int, const char*, int, const char*) TODO("mc_env_multi")
extern "C" int mc_table_print_kmers(const mc_table*, int, const char*, const char*, uint64_t*) TODO("mc_table_print_kmers")
extern "C" int mc_table_load_kmers(mc_table*, const char*, int) TODO("mc_table_load_kmers")
