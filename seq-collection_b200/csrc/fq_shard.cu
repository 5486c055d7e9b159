// fq_shard.cu -- multi-GPU byte-range shard protocol (include/fqgpu.h).  Placeholder
// until the single-GPU path is measured; every entry point fails loudly.
#include "fq_layout.h"
extern "C" {
size_t fqgpu_shard_block_words(void) { return (size_t)fq::BLOCK_WORDS + 32; }
int fqgpu_shard_begin(fqgpu_ctx*, int, int) { return FQGPU_EARG; }
int fqgpu_shard_export(fqgpu_ctx*, uint64_t*) { return FQGPU_EARG; }
int fqgpu_shard_combine(fqgpu_ctx*, const uint64_t*, fqgpu_stats*) { return FQGPU_EARG; }
int fqgpu_shard_rescan(fqgpu_ctx*, const uint64_t*) { return FQGPU_EARG; }
}
