#include "common.cuh"
namespace shrimp { void free_genome(shrimp_gpu_ctx*) {} void free_pipeline(shrimp_gpu_ctx*) {} }
