// hm_piz.cpp — decoder for OpenEXR PIZ-compressed scanline blocks.
//
// The reference loads its environment maps through tinyexr (model.cpp:158-231, loadEnvTexture);
// both shipped maps (scenes/envmaps/*.exr) are PIZ-compressed, so the scene loader needs this
// to accept the shipped config.json files unchanged.  Written from the OpenEXR file-format
// description (PIZ = value-range compaction through a bitmap LUT, a 2-D Haar-style wavelet on
// 16-bit words, canonical Huffman coding with run-length escapes); load-time CPU work only.
//
// Block layout after the chunk header:
//   u16 minNonZero, u16 maxNonZero, bitmap bytes [minNonZero .. maxNonZero]
//   i32 huffman byte count, then the Huffman stream:
//       u32 im, u32 iM, u32 table bytes, u32 nBits, u32 unused, packed code lengths, code bits
//   decoded words are channel-planar: per channel ny rows of nx * (bytes per sample / 2) words
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace hm {

namespace {

constexpr int kEncBits = 16;
constexpr int kEncSize = (1 << kEncBits) + 1;   // symbols 0..65535 plus the run-length escape
constexpr int kMaxCodeLen = 58;
constexpr int kFastBits = 12;

struct BitReader {
    const uint8_t* p;
    size_t nbits;      // total bits available
    size_t pos = 0;    // next bit, MSB-first within each byte
    // next `n` bits (n <= 32) without consuming; bits past the end read as 0
    uint32_t peek(int n) const {
        uint64_t v = 0;
        size_t byte = pos >> 3;
        const size_t nbytes = (nbits + 7) >> 3;
        for (int i = 0; i < 6; ++i) v = (v << 8) | (byte + i < nbytes ? p[byte + i] : 0);
        const int shift = 48 - (int)(pos & 7) - n;
        return (uint32_t)((v >> shift) & ((1ull << n) - 1));
    }
    void skip(int n) { pos += (size_t)n; }
    uint32_t get(int n) { uint32_t v = peek(n); pos += (size_t)n; return v; }
    bool exhausted() const { return pos >= nbits; }
};

[[noreturn]] void bad(const char* what) { throw std::invalid_argument(std::string("EXR PIZ block is corrupt: ") + what); }

uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

// Canonical Huffman decoder over code lengths (longest codes get the numerically smallest values).
struct Huffman {
    uint64_t base[kMaxCodeLen + 2];      // first code of each length
    uint32_t count[kMaxCodeLen + 2];
    uint32_t offset[kMaxCodeLen + 2];    // into `symbols`
    std::vector<uint32_t> symbols;       // sorted by (length, symbol)
    std::vector<uint32_t> fast;          // kFastBits lookup: symbol << 6 | length, 0 = long code
    int max_len = 0;

    void build(const std::vector<uint8_t>& len) {
        memset(count, 0, sizeof(count));
        for (size_t i = 0; i < len.size(); ++i) count[len[i]]++;
        count[0] = 0;
        uint64_t c = 0;
        for (int l = kMaxCodeLen; l > 0; --l) {
            const uint64_t next = (c + count[l]) >> 1;
            base[l] = c;
            c = next;
            if (count[l] && l > max_len) max_len = l;
        }
        uint32_t o = 0;
        for (int l = 1; l <= kMaxCodeLen; ++l) { offset[l] = o; o += count[l]; }
        symbols.resize(o);
        std::vector<uint32_t> fill(kMaxCodeLen + 2, 0);
        for (size_t i = 0; i < len.size(); ++i)
            if (len[i]) symbols[offset[len[i]] + fill[len[i]]++] = (uint32_t)i;
        fast.assign((size_t)1 << kFastBits, 0);
        for (int l = 1; l <= kFastBits && l <= max_len; ++l)
            for (uint32_t k = 0; k < count[l]; ++k) {
                const uint64_t code = base[l] + k;
                const uint32_t sym = symbols[offset[l] + k];
                const uint32_t first = (uint32_t)(code << (kFastBits - l));
                for (uint32_t j = 0; j < (1u << (kFastBits - l)); ++j) fast[first + j] = (sym << 6) | (uint32_t)l;
            }
    }

    // returns the symbol, or -1 when the bits do not form a code
    int decode(BitReader& br) const {
        const uint32_t e = fast[br.peek(kFastBits)];
        if (e) { br.skip((int)(e & 63)); return (int)(e >> 6); }
        // longer than the lookup covers: walk the canonical ranges bit by bit
        const size_t start = br.pos;
        uint64_t code = 0;
        for (int l = 1; l <= max_len; ++l) {
            code = (code << 1) | br.get(1);
            if (code >= base[l] && code - base[l] < count[l]) return (int)symbols[offset[l] + (code - base[l])];
        }
        br.pos = start;
        return -1;
    }
};

void huf_uncompress(const uint8_t* src, size_t n, uint16_t* out, size_t n_out) {
    if (n == 0) { if (n_out) bad("empty Huffman stream"); return; }
    if (n < 20) bad("short Huffman header");
    const uint32_t im = rd32(src), iM = rd32(src + 4), nbits = rd32(src + 12);
    if (im >= (uint32_t)kEncSize || iM >= (uint32_t)kEncSize || im > iM) bad("symbol range");
    const uint8_t* p = src + 20;
    const size_t avail = n - 20;

    // code lengths: 6 bits each, with zero-run escapes (59..62: 2..5 zeros; 63: 8-bit count + 6)
    std::vector<uint8_t> len(kEncSize, 0);
    BitReader tr{p, avail * 8};
    for (uint32_t s = im; s <= iM; ++s) {
        if (tr.pos + 6 > tr.nbits) bad("code-length table");
        const uint32_t l = tr.get(6);
        if (l == 63) {
            if (tr.pos + 8 > tr.nbits) bad("code-length table");
            uint32_t run = tr.get(8) + 6;
            if (s + run > iM + 1) bad("zero run past the table");
            s += run - 1;
        } else if (l >= 59) {
            uint32_t run = l - 59 + 2;
            if (s + run > iM + 1) bad("zero run past the table");
            s += run - 1;
        } else {
            len[s] = (uint8_t)l;
        }
    }
    const size_t table_bytes = (tr.pos + 7) >> 3;
    if ((size_t)nbits > (avail - table_bytes) * 8) bad("bit count");

    Huffman h;
    h.build(len);
    BitReader br{p + table_bytes, nbits};
    size_t o = 0;
    const int rlc = (int)iM;   // run-length escape: repeat the previous word `next byte` times
    while (o < n_out) {
        if (br.exhausted()) bad("not enough data");
        const int sym = h.decode(br);
        if (sym < 0) bad("invalid code");
        if (sym == rlc) {
            const uint32_t run = br.get(8);
            if (o == 0 || o + run > n_out) bad("run length");
            const uint16_t v = out[o - 1];
            for (uint32_t k = 0; k < run; ++k) out[o++] = v;
        } else {
            out[o++] = (uint16_t)sym;
        }
    }
}

// inverse of the wavelet's butterfly, 14-bit flavour (data range < 2^14: plain integer arithmetic)
inline void wdec14(uint16_t l, uint16_t h, uint16_t& a, uint16_t& b) {
    const int ls = (int16_t)l, hs = (int16_t)h;
    const int ai = ls + (hs & 1) + (hs >> 1);
    a = (uint16_t)(int16_t)ai;
    b = (uint16_t)(int16_t)(ai - hs);
}
// 16-bit flavour: modulo arithmetic around an offset of 2^15
inline void wdec16(uint16_t l, uint16_t h, uint16_t& a, uint16_t& b) {
    const int m = l, d = h;
    const int bb = (m - (d >> 1)) & 0xffff;
    const int aa = (d + bb - 0x8000) & 0xffff;
    b = (uint16_t)bb;
    a = (uint16_t)aa;
}

// in: first word of the plane; nx words per row with stride ox, ny rows with stride oy
void wav2_decode(uint16_t* in, int nx, int ox, int ny, int oy, uint16_t mx) {
    const bool w14 = mx < (1 << 14);
    const int n = nx > ny ? ny : nx;
    int p = 1;
    while (p <= n) p <<= 1;
    p >>= 1;
    int p2 = p;
    p >>= 1;
    while (p >= 1) {
        uint16_t* py = in;
        uint16_t* const ey = in + (ptrdiff_t)oy * (ny - p2);
        const ptrdiff_t oy1 = (ptrdiff_t)oy * p, oy2 = (ptrdiff_t)oy * p2, ox1 = (ptrdiff_t)ox * p, ox2 = (ptrdiff_t)ox * p2;
        uint16_t i00, i01, i10, i11;
        for (; py <= ey; py += oy2) {
            uint16_t* px = py;
            uint16_t* const ex = py + (ptrdiff_t)ox * (nx - p2);
            for (; px <= ex; px += ox2) {
                uint16_t* p01 = px + ox1; uint16_t* p10 = px + oy1; uint16_t* p11 = p10 + ox1;
                if (w14) {
                    wdec14(*px, *p10, i00, i10); wdec14(*p01, *p11, i01, i11);
                    wdec14(i00, i01, *px, *p01); wdec14(i10, i11, *p10, *p11);
                } else {
                    wdec16(*px, *p10, i00, i10); wdec16(*p01, *p11, i01, i11);
                    wdec16(i00, i01, *px, *p01); wdec16(i10, i11, *p10, *p11);
                }
            }
            if (nx & p) {   // odd column left over at this level
                uint16_t* p10 = px + oy1;
                if (w14) wdec14(*px, *p10, i00, *p10); else wdec16(*px, *p10, i00, *p10);
                *px = i00;
            }
        }
        if (ny & p) {       // odd row left over at this level
            uint16_t* px = py;
            uint16_t* const ex = py + (ptrdiff_t)ox * (nx - p2);
            for (; px <= ex; px += ox2) {
                uint16_t* p01 = px + ox1;
                if (w14) wdec14(*px, *p01, i00, *p01); else wdec16(*px, *p01, i00, *p01);
                *px = i00;
            }
        }
        p2 = p;
        p >>= 1;
    }
}

}  // namespace

// out: the block in the uncompressed scanline layout (per row: per channel nx samples), as
// 16-bit words; chan_u16_per_pixel[c] = 1 for HALF, 2 for FLOAT/UINT channels.
void piz_decompress(const uint8_t* src, size_t src_len, uint16_t* out, size_t out_count,
                    const std::vector<int>& chan_u16_per_pixel, int nx, int ny) {
    if (src_len < 4) bad("short block");
    const uint16_t min_nz = (uint16_t)(src[0] | (src[1] << 8)), max_nz = (uint16_t)(src[2] | (src[3] << 8));
    size_t p = 4;
    std::vector<uint8_t> bitmap(8192, 0);
    if (min_nz <= max_nz) {
        if (max_nz >= 8192) bad("bitmap range");
        const size_t nb = (size_t)max_nz - min_nz + 1;
        if (p + nb > src_len) bad("bitmap");
        memcpy(bitmap.data() + min_nz, src + p, nb);
        p += nb;
    }
    std::vector<uint16_t> lut(65536, 0);
    int k = 0;
    for (int i = 0; i < 65536; ++i)
        if (i == 0 || (bitmap[i >> 3] & (1 << (i & 7)))) lut[k++] = (uint16_t)i;
    const uint16_t max_value = (uint16_t)(k - 1);

    if (p + 4 > src_len) bad("Huffman length");
    const uint32_t hlen = rd32(src + p);
    p += 4;
    if (p + hlen > src_len) bad("Huffman length");

    size_t total = 0;
    for (int s : chan_u16_per_pixel) total += (size_t)s * nx * ny;
    if (total != out_count) bad("size mismatch");
    std::vector<uint16_t> tmp(total);
    huf_uncompress(src + p, hlen, tmp.data(), total);

    // wavelet per channel (and per 16-bit half of 32-bit samples), then the LUT
    size_t start = 0;
    std::vector<size_t> chan_start(chan_u16_per_pixel.size());
    for (size_t c = 0; c < chan_u16_per_pixel.size(); ++c) {
        const int size = chan_u16_per_pixel[c];
        chan_start[c] = start;
        for (int j = 0; j < size; ++j) wav2_decode(tmp.data() + start + j, nx, size, ny, nx * size, max_value);
        start += (size_t)size * nx * ny;
    }
    for (size_t i = 0; i < total; ++i) tmp[i] = lut[tmp[i]];

    // channel-planar -> scanline-interleaved
    uint16_t* dst = out;
    std::vector<size_t> cursor = chan_start;
    for (int y = 0; y < ny; ++y)
        for (size_t c = 0; c < chan_u16_per_pixel.size(); ++c) {
            const size_t n = (size_t)chan_u16_per_pixel[c] * nx;
            memcpy(dst, tmp.data() + cursor[c], n * 2);
            dst += n;
            cursor[c] += n;
        }
}

}  // namespace hm
