// bf16 attention entry point.  Bring-up version: routes to the SIMT kernel on bf16 inputs;
// replaced by the tcgen05 flash-attention kernel.
#include "common.cuh"

namespace pdm {
void attention_tc_bf16(const bf16* qkv, bf16* out, int nb, int L, int H, cudaStream_t s) {
    attention_simt(qkv, out, nb, L, H, true, s);
}
}  // namespace pdm
