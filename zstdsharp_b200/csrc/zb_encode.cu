// zb_encode.cu -- placeholder until the encoder kernels land (next commit).
#include "zb_encode.cuh"
namespace zb {
struct EncArenaImpl {};
void EncArena::release() {}
const uint8_t* EncArena::compactBuf() const { return nullptr; }
static const char* g_err = "encoder not built yet";
bool enc_compress_device(EncArena&, cudaStream_t, cudaEvent_t*, size_t, int, const uint8_t*, const uint64_t*, const size_t*, uint8_t*, const uint64_t*, const size_t*, size_t*, float*, unsigned*) { return false; }
bool enc_compact_device(EncArena&, cudaStream_t, size_t, const uint8_t*, const uint64_t*, const size_t*, const uint64_t*, size_t, unsigned*) { return false; }
const char* enc_last_error() { return g_err; }
}
