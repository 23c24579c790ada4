// PNG reading/writing on zlib (no libpng in this toolchain).  Host-side, off the render path.
//
// Writer contract (reference src/image/image_io.cpp:109-152): non-interlaced 8-bit RGBA, each channel
// round(255 * v) clamped to [0, 255].  Reader contract (:31-85): 8-bit images are expanded to RGBA and mapped to
// value / 255; colour types grey, grey+alpha, RGB, RGBA and palette are accepted.  Failures throw std::logic_error.
#include <PathTrace/image/image_io.h>

#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iterator>
#include <stdexcept>
#include <vector>

namespace {

    constexpr unsigned char kSignature[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};

    void putBE32(std::vector<unsigned char> &out, uint32_t v) {
        out.push_back(static_cast<unsigned char>(v >> 24));
        out.push_back(static_cast<unsigned char>(v >> 16));
        out.push_back(static_cast<unsigned char>(v >> 8));
        out.push_back(static_cast<unsigned char>(v));
    }

    uint32_t getBE32(const unsigned char *p) {
        return (static_cast<uint32_t>(p[0]) << 24) | (static_cast<uint32_t>(p[1]) << 16) | (static_cast<uint32_t>(p[2]) << 8) | p[3];
    }

    void putChunk(std::vector<unsigned char> &out, const char type[4], const unsigned char *payload, std::size_t length) {
        putBE32(out, static_cast<uint32_t>(length));
        const std::size_t type_at = out.size();
        out.insert(out.end(), type, type + 4);
        out.insert(out.end(), payload, payload + length);
        const uint32_t crc = static_cast<uint32_t>(crc32(0L, out.data() + type_at, static_cast<uInt>(4 + length)));
        putBE32(out, crc);
    }

    // reference src/image/image_io.cpp:139-142: round(255.0 * v) evaluated in DOUBLE, then clamped to [0, 255] (a float
    // product rounds some values across the .5 boundary: v = 0x1.383838p-1 gives 156 in float, 155 in double)
    unsigned char quantise(float v) {
        const double scaled = std::round(255.0 * static_cast<double>(v));
        if(!(scaled > 0.0)) {
            return 0;
        }
        return scaled >= 255.0 ? 255 : static_cast<unsigned char>(scaled);
    }

    int paeth(int a, int b, int c) {
        const int p = a + b - c;
        const int pa = std::abs(p - a);
        const int pb = std::abs(p - b);
        const int pc = std::abs(p - c);
        if(pa <= pb && pa <= pc) {
            return a;
        }
        return pb <= pc ? b : c;
    }

}

namespace io {

    void writeRGBImage(std::basic_ostream<char> &stream, const Image<Color<float>> &image) noexcept(false) {
        const int width = image.getWidth();
        const int height = image.getHeight();
        if(width <= 0 || height <= 0) {
            throw std::logic_error("writeRGBImage: empty image");
        }

        // filter type 0 scanlines
        std::vector<unsigned char> raw(static_cast<std::size_t>(height) * (1 + 4 * static_cast<std::size_t>(width)));
        std::size_t w = 0;
        for(int y = 0; y < height; y++) {
            raw[w++] = 0;
            for(int x = 0; x < width; x++) {
                const Color<float> pixel = image(x, y);
                for(int channel = 0; channel < 4; channel++) {
                    raw[w++] = quantise(pixel[channel]);
                }
            }
        }

        uLongf packed_size = compressBound(static_cast<uLong>(raw.size()));
        std::vector<unsigned char> packed(packed_size);
        if(compress2(packed.data(), &packed_size, raw.data(), static_cast<uLong>(raw.size()), Z_DEFAULT_COMPRESSION) != Z_OK) {
            throw std::logic_error("writeRGBImage: deflate failed");
        }

        std::vector<unsigned char> out(kSignature, kSignature + 8);
        std::vector<unsigned char> header;
        putBE32(header, static_cast<uint32_t>(width));
        putBE32(header, static_cast<uint32_t>(height));
        header.push_back(8); // bit depth
        header.push_back(6); // RGBA
        header.push_back(0); // deflate
        header.push_back(0); // adaptive filtering
        header.push_back(0); // no interlace
        putChunk(out, "IHDR", header.data(), header.size());
        putChunk(out, "IDAT", packed.data(), packed_size);
        putChunk(out, "IEND", nullptr, 0);

        stream.write(reinterpret_cast<const char *>(out.data()), static_cast<std::streamsize>(out.size()));
        if(!stream) {
            throw std::logic_error("writeRGBImage: stream write failed");
        }
    }

    void writeRGBImage(const std::string &path, const Image<Color<float>> &image) noexcept(false) {
        std::ofstream stream(path, std::ios_base::out | std::ios_base::binary);
        if(!stream) {
            throw std::logic_error("writeRGBImage: cannot open " + path);
        }
        writeRGBImage(stream, image);
    }

    void writeRGBImage(const std::filesystem::path &path, const Image<Color<float>> &image) noexcept(false) {
        writeRGBImage(path.string(), image);
    }

    namespace {
        Image<Color<float>> decodePng(std::basic_istream<char> &stream);
    }

    // Reads every PNG the reference reads through libpng with EXPAND | PACKING | STRIP_16 (image_io.cpp:50-107): all five
    // colour types, bit depths 1 / 2 / 4 / 8 / 16, non-interlaced and Adam7.  What those transforms leave is an 8-bit image
    // of 1 (grey), 2 (grey + alpha), 3 (RGB) or 4 (RGBA) channels -- palette entries become RGB(A), a tRNS colour key
    // becomes an alpha channel, grey samples below 8 bits are scaled to 8, 16-bit samples keep their high byte -- and the
    // reference then copies 3- and 4-channel rows only (image_io.cpp:63-79): a grey or grey + alpha file yields an image of
    // zero-initialised pixels.  That behaviour is kept.  Every failure, including running out of memory on a hostile
    // header, is a std::logic_error.
    Image<Color<float>> readRGBImage(std::basic_istream<char> &stream) noexcept(false) {
        try {
            return decodePng(stream);
        }
        catch(const std::bad_alloc &) {
            throw std::logic_error("readRGBImage: out of memory while decoding");
        }
    }

    namespace {
    Image<Color<float>> decodePng(std::basic_istream<char> &stream) {
        const std::vector<unsigned char> file((std::istreambuf_iterator<char>(stream)), std::istreambuf_iterator<char>());
        if(file.size() < 8 || std::memcmp(file.data(), kSignature, 8) != 0) {
            throw std::logic_error("readRGBImage: not a PNG stream");
        }

        uint32_t width = 0;
        uint32_t height = 0;
        int bit_depth = 0;
        int colour_type = -1;
        int interlace = 0;
        std::vector<unsigned char> palette;
        std::vector<unsigned char> transparency;
        std::vector<unsigned char> packed;
        bool ended = false;

        std::size_t at = 8;
        while(!ended && at + 12 <= file.size()) {
            const uint32_t length = getBE32(&file[at]);
            if(at + 12 + static_cast<std::size_t>(length) > file.size()) {
                throw std::logic_error("readRGBImage: truncated chunk");
            }
            const unsigned char *type = &file[at + 4];
            const unsigned char *payload = &file[at + 8];
            const uint32_t crc = getBE32(&file[at + 8 + length]);
            if(crc != static_cast<uint32_t>(crc32(0L, type, static_cast<uInt>(4 + length)))) {
                throw std::logic_error("readRGBImage: chunk checksum mismatch");
            }
            if(std::memcmp(type, "IHDR", 4) == 0) {
                if(length != 13) {
                    throw std::logic_error("readRGBImage: bad header");
                }
                width = getBE32(payload);
                height = getBE32(payload + 4);
                bit_depth = payload[8];
                colour_type = payload[9];
                interlace = payload[12];
            }
            else if(std::memcmp(type, "PLTE", 4) == 0) {
                palette.assign(payload, payload + length);
            }
            else if(std::memcmp(type, "tRNS", 4) == 0) {
                transparency.assign(payload, payload + length);
            }
            else if(std::memcmp(type, "IDAT", 4) == 0) {
                packed.insert(packed.end(), payload, payload + length);
            }
            else if(std::memcmp(type, "IEND", 4) == 0) {
                ended = true;
            }
            at += 12 + static_cast<std::size_t>(length);
        }

        int channels = 0;
        switch(colour_type) {
            case 0:
                channels = 1;
                break;
            case 2:
                channels = 3;
                break;
            case 3:
                channels = 1;
                break;
            case 4:
                channels = 2;
                break;
            case 6:
                channels = 4;
                break;
            default:
                throw std::logic_error("readRGBImage: unsupported colour type");
        }
        const bool depth_ok = colour_type == 0   ? (bit_depth == 1 || bit_depth == 2 || bit_depth == 4 || bit_depth == 8 || bit_depth == 16)
                              : colour_type == 3 ? (bit_depth == 1 || bit_depth == 2 || bit_depth == 4 || bit_depth == 8)
                                                 : (bit_depth == 8 || bit_depth == 16);
        if(width == 0 || height == 0 || width > 65535U || height > 65535U || !depth_ok || interlace > 1) {
            throw std::logic_error("readRGBImage: invalid header");
        }

        // the passes of the image: one for a non-interlaced file, seven for Adam7 (x0, y0, dx, dy per pass)
        struct Pass {
            uint32_t x0, y0, dx, dy;
        };
        static const Pass kAdam7[7] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};
        static const Pass kWhole[1] = {{0, 0, 1, 1}};
        const Pass *passes = interlace != 0 ? kAdam7 : kWhole;
        const int n_passes = interlace != 0 ? 7 : 1;
        const std::size_t bits_per_pixel = static_cast<std::size_t>(channels) * static_cast<std::size_t>(bit_depth);
        const std::size_t filter_unit = std::max<std::size_t>(bits_per_pixel / 8, 1); // bytes per complete pixel, at least 1

        std::size_t raw_bytes = 0;
        for(int k = 0; k < n_passes; k++) {
            const Pass &ps = passes[k];
            const std::size_t pw = width > ps.x0 ? (width - ps.x0 + ps.dx - 1) / ps.dx : 0;
            const std::size_t ph = height > ps.y0 ? (height - ps.y0 + ps.dy - 1) / ps.dy : 0;
            if(pw != 0 && ph != 0) {
                raw_bytes += ph * (1 + (pw * bits_per_pixel + 7) / 8);
            }
        }
        // deflate expands at most ~1032:1: an IHDR that promises more pixels than the IDAT stream can possibly hold is
        // rejected before anything is allocated (a 100-byte file must not trigger a multi-gigabyte allocation)
        if(static_cast<long double>(raw_bytes) > 1040.0L * static_cast<long double>(packed.size()) + 1024.0L) {
            throw std::logic_error("readRGBImage: image dimensions exceed what the compressed data can hold");
        }
        std::vector<unsigned char> raw(raw_bytes);
        uLongf raw_size = static_cast<uLongf>(raw.size());
        if(uncompress(raw.data(), &raw_size, packed.data(), static_cast<uLong>(packed.size())) != Z_OK || raw_size != raw.size()) {
            throw std::logic_error("readRGBImage: inflate failed");
        }

        // samples[(y * width + x) * channels + c]: the file's samples, one per element (PACKING: values below 8 bits unscaled)
        std::vector<uint16_t> samples(static_cast<std::size_t>(width) * height * channels);
        std::vector<unsigned char> line;
        std::vector<unsigned char> previous;
        std::size_t in = 0;
        for(int k = 0; k < n_passes; k++) {
            const Pass &ps = passes[k];
            const std::size_t pw = width > ps.x0 ? (width - ps.x0 + ps.dx - 1) / ps.dx : 0;
            const std::size_t ph = height > ps.y0 ? (height - ps.y0 + ps.dy - 1) / ps.dy : 0;
            if(pw == 0 || ph == 0) {
                continue;
            }
            const std::size_t line_bytes = (pw * bits_per_pixel + 7) / 8;
            line.assign(line_bytes, 0);
            previous.assign(line_bytes, 0);
            for(std::size_t row = 0; row < ph; row++) {
                const unsigned char filter = raw[in++];
                for(std::size_t i = 0; i < line_bytes; i++) {
                    const int left = i >= filter_unit ? line[i - filter_unit] : 0;
                    const int up = previous[i];
                    const int up_left = i >= filter_unit ? previous[i - filter_unit] : 0;
                    int predicted = 0;
                    switch(filter) {
                        case 0:
                            predicted = 0;
                            break;
                        case 1:
                            predicted = left;
                            break;
                        case 2:
                            predicted = up;
                            break;
                        case 3:
                            predicted = (left + up) / 2;
                            break;
                        case 4:
                            predicted = paeth(left, up, up_left);
                            break;
                        default:
                            throw std::logic_error("readRGBImage: unknown scanline filter");
                    }
                    line[i] = static_cast<unsigned char>(raw[in + i] + predicted);
                }
                in += line_bytes;
                const std::size_t y = ps.y0 + row * ps.dy;
                for(std::size_t col = 0; col < pw; col++) {
                    const std::size_t x = ps.x0 + col * ps.dx;
                    uint16_t *out = &samples[(y * width + x) * channels];
                    for(int c = 0; c < channels; c++) {
                        const std::size_t sample = col * channels + static_cast<std::size_t>(c);
                        if(bit_depth == 16) {
                            out[c] = static_cast<uint16_t>((static_cast<unsigned>(line[2 * sample]) << 8) | line[2 * sample + 1]);
                        }
                        else if(bit_depth == 8) {
                            out[c] = line[sample];
                        }
                        else {
                            const std::size_t bit = sample * static_cast<std::size_t>(bit_depth);
                            const int shift = 8 - bit_depth - static_cast<int>(bit % 8);
                            out[c] = static_cast<uint16_t>((line[bit / 8] >> shift) & ((1 << bit_depth) - 1));
                        }
                    }
                }
                previous.swap(line);
            }
        }

        // EXPAND: palette -> RGB(A); a tRNS colour key -> alpha; grey below 8 bits scaled to 8 (libpng: value * 255 / max)
        Image<Color<float>> image(static_cast<int>(width), static_cast<int>(height));
        const bool keyed = !transparency.empty() && (colour_type == 0 || colour_type == 2);
        const int out_channels = colour_type == 3 ? (transparency.empty() ? 3 : 4) : channels + (keyed ? 1 : 0);
        if(out_channels != 3 && out_channels != 4) {
            return image; // grey and grey + alpha: the reference copies no rows (image_io.cpp:63-79)
        }
        // the colour key is compared with the file's samples at their full depth (libpng expands before it strips), then
        // STRIP_16 keeps the high byte
        auto key = [&](int c) -> unsigned {
            return transparency.size() >= static_cast<std::size_t>(2 * c + 2) ? (static_cast<unsigned>(transparency[2 * c]) << 8) | transparency[2 * c + 1] : 0x10000U;
        };
        const int strip = bit_depth == 16 ? 8 : 0;
        for(uint32_t y = 0; y < height; y++) {
            for(uint32_t x = 0; x < width; x++) {
                const uint16_t *p = &samples[(static_cast<std::size_t>(y) * width + x) * channels];
                unsigned char rgba[4] = {0, 0, 0, 255};
                if(colour_type == 3) {
                    const std::size_t entry = p[0];
                    if(3 * entry + 2 >= palette.size()) {
                        throw std::logic_error("readRGBImage: palette index out of range");
                    }
                    rgba[0] = palette[3 * entry];
                    rgba[1] = palette[3 * entry + 1];
                    rgba[2] = palette[3 * entry + 2];
                    if(entry < transparency.size()) {
                        rgba[3] = transparency[entry];
                    }
                }
                else if(colour_type == 2) {
                    rgba[0] = static_cast<unsigned char>(p[0] >> strip);
                    rgba[1] = static_cast<unsigned char>(p[1] >> strip);
                    rgba[2] = static_cast<unsigned char>(p[2] >> strip);
                    if(keyed && p[0] == key(0) && p[1] == key(1) && p[2] == key(2)) {
                        rgba[3] = 0;
                    }
                }
                else {
                    for(int c = 0; c < 4; c++) {
                        rgba[c] = static_cast<unsigned char>(p[c] >> strip);
                    }
                }
                image(static_cast<int>(x), static_cast<int>(y)) =
                  Color<float>(rgba[0] / 255.0F, rgba[1] / 255.0F, rgba[2] / 255.0F, rgba[3] / 255.0F);
            }
        }
        return image;
    }

    } // namespace

    Image<Color<float>> readRGBImage(const std::string &path) noexcept(false) {
        std::ifstream stream(path, std::ios_base::in | std::ios_base::binary);
        if(!stream) {
            throw std::logic_error("readRGBImage: cannot open " + path);
        }
        return readRGBImage(stream);
    }

    Image<Color<float>> readRGBImage(const std::filesystem::path &path) noexcept(false) {
        return readRGBImage(path.string());
    }

}
