// hm_piz.cpp — OpenEXR PIZ block decoder (placeholder until the wavelet/Huffman decoder lands).
#include <stdexcept>
#include <vector>
#include <cstdint>
#include <cstddef>
namespace hm {
void piz_decompress(const uint8_t*, size_t, uint16_t*, size_t, const std::vector<int>&, int, int) {
    throw std::invalid_argument("EXR PIZ compression is not supported yet");
}
}
