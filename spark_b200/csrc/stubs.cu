// stubs.cu -- entry points declared in include/spark_b200.h whose kernels are not built yet.
// They fail loudly (SB_ERR_UNSUPPORTED); there is no CPU fallback behind any of them.
#include "common.cuh"
using namespace sb;
#define SB_STUB(name, ...) \
  extern "C" int name(__VA_ARGS__) { set_last_error(#name " is not implemented in this build"); return SB_ERR_UNSUPPORTED; }
SB_STUB(sb_join_build, const sb_table *, const int32_t *, int32_t, sb_stream *, sb_hash_table **)
SB_STUB(sb_join_probe, const sb_hash_table *, const sb_table *, const int32_t *, int32_t, int32_t, sb_stream *, sb_table **)
SB_STUB(sb_hash_table_release, sb_hash_table *)
